"""Host logic of the POSIX-shared-memory persistent wavefunction (qvm_b200/shm.py): object layout and info-socket protocol of the
reference's `--shared` mode (src/shm.lisp:181-229, src/impl/sbcl.lisp:9-10,40-53, app/src/impl/sbcl.lisp:10-42), with a host array
standing in for the device state (the GPU variant is in tests/test_gpu_boundary.py)."""
import mmap
import os
import struct
import uuid

import numpy as np
import pytest

from qvm_b200 import shm


def _name():
    return f"QVMTEST{uuid.uuid4().hex[:12]}"


def test_shared_object_layout_and_info_socket(tmp_path):
    n = 10
    state = (np.arange(1 << n) + 1j * np.arange(1 << n)[::-1]).astype(np.complex128)
    name = _name()

    def download(dst):
        dst[:] = state

    def upload(src):
        state[:] = src

    with shm.SharedWavefunction(name, 1 << n, download, upload, socket_dir=str(tmp_path)) as sw:
        # fresh object: |0...0> (make-shared-wavefunction), header = (widetag, fixnum length)
        assert sw.amplitudes[0] == 1.0 and not sw.amplitudes[1:].any()
        assert os.path.getsize(f"/dev/shm/{name}") % mmap.PAGESIZE == 0
        assert os.path.getsize(f"/dev/shm/{name}") >= shm.HEADER_BYTES + 16 * (1 << n)
        with open(f"/dev/shm/{name}", "rb") as f:
            w0, w1 = struct.unpack("<QQ", f.read(16))
        assert (w0, w1) == (0, (1 << n) << 1)
        # the info socket answers "<length>,<offset>" to any client that sends one octet; more than once
        for _ in range(3):
            assert shm.query_info(name, str(tmp_path)) == (1 << n, shm.HEADER_BYTES)
        sw.refresh()
        view = shm.attach(name, str(tmp_path))          # what another process does
        assert np.array_equal(view, state)
        view[5] = 42.0                                   # a client writes; the owner pushes it to the device
        sw.push()
        assert state[5] == 42.0
        # a name in use is an error (O_EXCL), as in the reference
        with pytest.raises(FileExistsError):
            shm.SharedWavefunction(name, 4, download, socket_dir=str(tmp_path))
        del view
    assert not os.path.exists(f"/dev/shm/{name}") and not os.path.exists(tmp_path / name)


def test_read_only_share_and_bad_names(tmp_path):
    with pytest.raises(ValueError):
        shm.SharedWavefunction("a/b", 4, lambda d: None, socket_dir=str(tmp_path))
    name = _name()
    with shm.SharedWavefunction(name, 4, lambda d: d.fill(0.5), socket_dir=str(tmp_path)) as sw:
        assert (sw.refresh() == 0.5).all()
        with pytest.raises(RuntimeError):
            sw.push()
