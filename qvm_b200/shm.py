"""POSIX-shared-memory persistent wavefunction (the app's `--shared NAME` mode).

Reference behaviour restated (app/src/entry-point.lisp:265-285, 513-526; src/shm.lisp:181-229, 280-300; src/impl/sbcl.lisp:9-10,
40-53; app/src/impl/sbcl.lisp:10-42): the wavefunction of a persistent QVM lives in a POSIX shared-memory object NAME as a Lisp
simple-array -- a two-word vector header (widetag, fixnum length) followed by 2^n (re, im) native doubles, the object rounded up to
whole pages -- and a local stream socket /tmp/NAME answers every connection, after reading one octet, with "<length>,<offset>" in
ASCII decimal, so that another process can map the object and find the amplitudes at byte OFFSET.

Here the amplitudes live in HBM; the shared object is the host-visible copy.  `refresh()` downloads the device state straight into
the mapped memory (qvmcuda_download writes there, no intermediate buffer), `push()` uploads what a client wrote.  The header words
carry the values SBCL x86-64 uses (simple-array (complex double-float) widetag 0xE5... is build-specific, so word 0 is written as the
caller-supplied widetag, default 0, and word 1 as the fixnum length = length << 1): clients of the reference only use OFFSET and LENGTH.
"""
from __future__ import annotations

import mmap
import os
import socket
import struct
import threading
from typing import Callable, Optional

import numpy as np

HEADER_BYTES = 16          # sb-vm:vector-data-offset (2) * n-word-bytes (8): shm-vector-header-size, src/impl/sbcl.lisp:9-10


def _round_to_page(size: int) -> int:
    page = mmap.PAGESIZE     # round-to-next-page, src/shm.lisp:177-179
    return (size + page - 1) // page * page


class SharedWavefunction:
    """The shared-memory object + its info socket.

    download(dst: np.ndarray[complex128]) fills dst with the current amplitudes; upload(src) is its inverse.  For a device state
    pass `vec.download_into` / `vec.upload` of a qvm_b200.qvm.DeviceVector (see `share_wavefunction`)."""

    def __init__(self, name: str, length: int, download: Callable[[np.ndarray], None],
                 upload: Optional[Callable[[np.ndarray], None]] = None, widetag: int = 0, socket_dir: str = "/tmp"):
        if not name or "/" in name:
            raise ValueError("shared memory name must be a non-empty string without '/'")
        self.name, self.length = name, int(length)
        self._download, self._upload = download, upload
        self.size = _round_to_page(HEADER_BYTES + 16 * self.length)
        # O_CREAT | O_EXCL | O_RDWR, mode rw-rw-rw- (make-posix-shared-memory, src/shm.lisp:188-199): a name in use is an error
        self._path = os.path.join("/dev/shm", name)
        fd = os.open(self._path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o666)
        try:
            os.ftruncate(fd, self.size)
            self._map = mmap.mmap(fd, self.size, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)          # "we don't need it after we've mmapped"
        struct.pack_into("<QQ", self._map, 0, widetag, self.length << 1)
        self.amplitudes = np.frombuffer(self._map, dtype=np.complex128, count=self.length, offset=HEADER_BYTES)
        self.amplitudes[:] = 0
        self.amplitudes[0] = 1.0       # make-shared-wavefunction: (setf (aref vec 0) (cflonum 1))
        self._sock_path = os.path.join(socket_dir, name)
        self._server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._server.bind(self._sock_path)
        self._server.listen(8)         # "8 is an arbitrary backlog value"
        self._closing = False
        self._thread = threading.Thread(target=self._serve, name=f"Socket server on {self._sock_path} for Shared Memory QVM",
                                        daemon=True)
        self._thread.start()

    # -- info server (start-shm-info-server, app/src/impl/sbcl.lisp:10-42)
    def _serve(self):
        response = f"{self.length},{HEADER_BYTES}".encode("ascii")
        while not self._closing:
            try:
                client, _ = self._server.accept()
            except OSError:
                break
            try:
                client.recv(1)
                client.sendall(response)
            except OSError:
                pass
            finally:
                client.close()

    # -- device <-> shared object
    def refresh(self) -> np.ndarray:
        """Device -> shared memory (what a client sees after a `run`)."""
        self._download(self.amplitudes)
        return self.amplitudes

    def push(self) -> None:
        """Shared memory -> device (a client wrote amplitudes)."""
        if self._upload is None:
            raise RuntimeError("this shared wavefunction is read-only")
        self._upload(self.amplitudes)

    def close(self) -> None:
        """free-posix-shared-memory (src/shm.lisp:232-249): unmap, unlink; the socket file is deleted with the server."""
        if self._closing:
            return
        self._closing = True
        try:
            self._server.close()
        finally:
            for p in (self._sock_path, self._path):
                try:
                    os.unlink(p)
                except FileNotFoundError:
                    pass
        self.amplitudes = None
        try:
            self._map.close()
        except BufferError:          # a client of this process still holds a view
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def query_info(name: str, socket_dir: str = "/tmp"):
    """What a client does: connect to /tmp/NAME, send one octet, read "<length>,<offset>"."""
    with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
        s.connect(os.path.join(socket_dir, name))
        s.sendall(b"?")
        data = b""
        while True:
            chunk = s.recv(64)
            if not chunk:
                break
            data += chunk
    length, offset = data.decode("ascii").split(",")
    return int(length), int(offset)


def attach(name: str, socket_dir: str = "/tmp") -> np.ndarray:
    """Client side: map the object read-write and return the amplitudes as a numpy view."""
    length, offset = query_info(name, socket_dir)
    fd = os.open(os.path.join("/dev/shm", name), os.O_RDWR)
    try:
        m = mmap.mmap(fd, 0, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
    finally:
        os.close(fd)
    return np.frombuffer(m, dtype=np.complex128, count=length, offset=offset)


def share_wavefunction(qvm, name: str, socket_dir: str = "/tmp") -> SharedWavefunction:
    """Persistent shared wavefunction of a PureStateQVM / DensityQVM: the device state exported through POSIX shared memory."""
    vec = qvm.state.vec
    length = vec.length

    def download(dst):
        vec.download_into(dst)

    def upload(src):
        vec.upload(src)

    shared = SharedWavefunction(name, length, download, upload, socket_dir=socket_dir)
    shared.refresh()
    return shared
