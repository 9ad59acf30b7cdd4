;;;; bindings.lisp -- CFFI view of include/qvmcuda.h (same conventions as src/shm.lisp:103-175 in the
;;;; reference: DEFINE-FOREIGN-LIBRARY / USE-FOREIGN-LIBRARY / DEFCFUN, C int status -> Lisp ERROR).

(in-package #:qvm-cuda)

(cffi:define-foreign-library libqvmcuda
  (t (:default "libqvmcuda")))

(cffi:use-foreign-library libqvmcuda)

(cffi:defcfun ("qvmcuda_last_error" %last-error) :string)

(defmacro defcuda (lisp-name c-name &rest args)
  "Define %LISP-NAME calling C-NAME and LISP-NAME signalling an ERROR on a non-zero status."
  (let ((raw (alexandria:symbolicate "%" lisp-name))
        (names (mapcar #'first args)))
    `(progn
       (cffi:defcfun (,c-name ,raw) :int ,@args)
       (defun ,lisp-name ,names
         (let ((status (,raw ,@names)))
           (unless (zerop status)
             (error "libqvmcuda: ~A failed: ~A" ,c-name (%last-error)))
           nil)))))

;;; allocation / state protocol
(defcuda state-create "qvmcuda_state_create" (n-amplitudes :uint64) (device :int) (out :pointer))
(defcuda state-destroy "qvmcuda_state_destroy" (state :pointer))
(defcuda synchronize "qvmcuda_synchronize" (state :pointer))
(defcuda download "qvmcuda_download" (state :pointer) (dst :pointer) (offset :uint64) (count :uint64))
(defcuda upload "qvmcuda_upload" (state :pointer) (src :pointer) (offset :uint64) (count :uint64))
(defcuda set-zero-state "qvmcuda_set_zero_state" (state :pointer))
(defcuda set-basis-state "qvmcuda_set_basis_state" (state :pointer) (basis :uint64))
(defcuda copy-state "qvmcuda_copy" (dst :pointer) (src :pointer))
;;; operator API
(defcuda apply-matrix "qvmcuda_apply_matrix" (state :pointer) (k :int) (qubits :pointer) (matrix :pointer))
(defcuda apply-gates "qvmcuda_apply_gates" (state :pointer) (n-gates :int) (ks :pointer) (qubits :pointer)
  (matrices :pointer) (flags :uint32))
(defcuda density-apply-kraus "qvmcuda_density_apply_kraus" (state :pointer) (n-qubits :int) (k :int)
  (qubits :pointer) (m :int) (kraus :pointer) (flags :uint32))
(defcuda density-apply-ops "qvmcuda_density_apply_ops" (state :pointer) (n-qubits :int) (n-ops :int) (ks :pointer)
  (qubits :pointer) (ms :pointer) (kraus :pointer) (flags :uint32))
;;; measurement protocol
(defcuda prob-excited "qvmcuda_prob_excited" (state :pointer) (qubit :int) (p :pointer))
(defcuda prob-ground "qvmcuda_prob_ground" (state :pointer) (qubit :int) (p :pointer))
(defcuda norm2 "qvmcuda_norm2" (state :pointer) (p :pointer))
(defcuda inner-product "qvmcuda_inner_product" (a :pointer) (b :pointer) (out :pointer))
(defcuda probabilities "qvmcuda_probabilities" (state :pointer) (out :pointer) (offset :uint64) (count :uint64))
(defcuda scale "qvmcuda_scale" (state :pointer) (factor :double))
(defcuda collapse "qvmcuda_collapse" (state :pointer) (qubit :int) (keep-bit :int) (inv-norm :double))
(defcuda sample "qvmcuda_sample" (state :pointer) (uniforms :pointer) (n-shots :uint64) (out :pointer) (strict :int))
(defcuda density-prob-excited "qvmcuda_density_prob_excited" (state :pointer) (n-qubits :int) (qubit :int) (p :pointer))
(defcuda density-collapse "qvmcuda_density_collapse" (state :pointer) (n-qubits :int) (qubit :int)
  (keep-bit :int) (inv-norm :double))
(defcuda density-measure-discard "qvmcuda_density_measure_discard" (state :pointer) (n-qubits :int) (qubit :int))
(defcuda density-diag-probs "qvmcuda_density_diag_probs" (state :pointer) (n-qubits :int) (out :pointer))

;;; expectation / unitary-qvm rows (app/src/api/expectation.lisp:78-107, src/unitary-qvm.lisp:30-137)
(defcuda density-expectation "qvmcuda_density_expectation" (state :pointer) (n-qubits :int) (op-matrix :pointer) (out :pointer))
(defcuda set-identity-matrix "qvmcuda_set_identity_matrix" (state :pointer) (n-qubits :int))
;;; several devices under one Lisp image (INTEGRATION.md section 3c): STATES is a foreign array of WORLD state pointers
(defcuda shard-attach-local "qvmcuda_shard_attach_local" (states :pointer) (world :int) (want-alt :int))
(defcuda shard-compile "qvmcuda_shard_compile" (state :pointer) (n-gates :int) (ks :pointer) (qubits :pointer)
  (matrices :pointer) (flags :uint32) (out :pointer))
(defcuda tape-num-steps "qvmcuda_tape_num_steps" (tape :pointer) (n-steps :pointer))
(defcuda tape-step-flags "qvmcuda_tape_step_flags" (tape :pointer) (step :int) (flags :pointer))
(defcuda tape-run-step "qvmcuda_tape_run_step" (state :pointer) (tape :pointer) (step :int))
(defcuda tape-commit "qvmcuda_tape_commit" (state :pointer) (tape :pointer))
(defcuda tape-destroy "qvmcuda_tape_destroy" (tape :pointer))
(defcuda state-layout "qvmcuda_state_layout" (state :pointer) (l2p :pointer) (n :int))
(defcuda shard-clear "qvmcuda_shard_clear" (state :pointer))
(defcuda shard-set-zero-ranks "qvmcuda_shard_set_zero_ranks" (state :pointer) (mask :uint32))

(defconstant +step-peer+ 1)
(defconstant +fuse+ 1)
(defconstant +absorb-swaps+ 2)

(defun call-returning-double (fn &rest args)
  (cffi:with-foreign-object (p :double)
    (apply fn (append args (list p)))
    (cffi:mem-ref p :double)))
