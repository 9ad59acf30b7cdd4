;;;; device-state.lisp -- a device-resident amplitude vector behind the reference's allocation and
;;;; state protocols (src/allocator.lisp:33-62, src/state-representation.lisp:25-47,58-167,173-286).
;;;;
;;;; Host visibility contract (SURVEY.md section 7): QVM::AMPLITUDES must hand out a real Lisp
;;;; (SIMPLE-ARRAY CFLONUM (*)) that callers AREF and even SETF.  The device vector therefore has a host
;;;; MIRROR: a foreign (static-vectors) array.  STATE-ELEMENTS downloads into it when the device copy is
;;;; newer; (SETF STATE-ELEMENTS) and RUN :BEFORE upload it when the host copy may be newer (it has been
;;;; handed out since the last upload -- raw AREF writes cannot be observed, so "handed out" = dirty).
;;;; With *CUDA-LAZY-MIRROR* = NIL the mirror is never allocated and AMPLITUDES signals an error: for
;;;; 30+ qubit states whose 16 GiB should not cross PCIe.

(in-package #:qvm-cuda)

(defvar *cuda-device* 0 "CUDA device ordinal new states are created on.")
(defvar *cuda-lazy-mirror* t "Keep a host mirror so that QVM::AMPLITUDES works (see file header).")

(defclass cuda-allocation ()
  ((length :initarg :length :reader qvm::allocation-length)
   (device :initarg :device :initform *cuda-device* :reader allocation-device))
  (:documentation "Allocation description for a device-resident vector.  Answers the two generics of
src/allocator.lisp:33-62; the hook for `--default-allocation cuda` (app/src/globals.lisp:17)."))

(defmethod qvm::allocate-vector ((descr cuda-allocation))
  "Returns the host MIRROR (zero-initialised, as the protocol requires) and a finalizer.  The device
vector itself is created by MAKE-DEVICE-PURE-STATE, which owns the handle."
  (let ((mirror (static-vectors:make-static-vector (qvm::allocation-length descr)
                                                   :element-type 'qvm:cflonum
                                                   :initial-element (qvm:cflonum 0))))
    (values mirror (lambda () (static-vectors:free-static-vector mirror) nil))))

(defclass device-state-mixin ()
  ((handle :initarg :handle :accessor device-handle :documentation "qvmcuda_state*")
   (device-newer-p :initform nil :accessor device-newer-p)
   (host-newer-p :initform nil :accessor host-newer-p)
   (tape :initform nil :accessor gate-tape
         :documentation "Pending gates (matrix . qubits), flushed in ONE qvmcuda_apply_gates call.")))

(defgeneric flush-gate-tape (state)
  (:documentation "Send the pending transitions of STATE to the GPU in one call (methods in operators.lisp).  Qubit lists go
out in NAT-TUPLE order (src/utilities.lisp:43-51), i.e. reversed Quil argument order."))

(defclass device-pure-state (device-state-mixin qvm::pure-state) ())
(defclass device-density-matrix-state (device-state-mixin qvm::density-matrix-state) ())

(defun %create-handle (length)
  (cffi:with-foreign-object (out :pointer)
    (state-create length *cuda-device* out)
    (cffi:mem-ref out :pointer)))

(defun %attach-finalizer (state handle mirror-finalizer)
  ;; finalizers run on arbitrary threads: qvmcuda_state_destroy is thread-safe and device-agnostic
  (tg:finalize state (lambda ()
                       (%state-destroy handle)
                       (when mirror-finalizer (funcall mirror-finalizer)))))

(defun make-device-pure-state (num-qubits)
  "MAKE-PURE-STATE (src/state-representation.lisp:76-102) with the amplitudes on the GPU."
  (let* ((length (expt 2 num-qubits))
         (handle (%create-handle length)))
    (set-zero-state handle)
    (multiple-value-bind (mirror finalizer)
        (if *cuda-lazy-mirror*
            (qvm::allocate-vector (make-instance 'cuda-allocation :length length))
            (values (qvm::make-lisp-cflonum-vector 2) nil)) ; placeholder, never read
      (setf (aref mirror 0) (qvm:cflonum 1))
      (let ((state (make-instance 'device-pure-state :num-qubits num-qubits
                                                     :amplitudes mirror
                                                     :handle handle)))
        (%attach-finalizer state handle finalizer)
        state))))

(defun make-device-density-matrix-state (num-qubits)
  "MAKE-DENSITY-MATRIX-STATE (src/state-representation.lisp:235-266) with vec(rho) on the GPU."
  (let* ((length (expt 2 (* 2 num-qubits)))
         (handle (%create-handle length)))
    (set-zero-state handle)
    (multiple-value-bind (mirror finalizer)
        (qvm::allocate-vector (make-instance 'cuda-allocation :length length))
      (setf (aref mirror 0) (qvm:cflonum 1))
      (let ((state (make-instance 'device-density-matrix-state :num-qubits num-qubits
                                                               :elements-vector mirror
                                                               :handle handle)))
        (%attach-finalizer state handle finalizer)
        state))))

;;; ---- mirror synchronisation ---------------------------------------------------------------

(defun %mirror (state)
  (etypecase state
    (device-pure-state (qvm::amplitudes state))
    (device-density-matrix-state (qvm::elements-vector state))))

(defun sync-to-host (state)
  (flush-gate-tape state)
  (when (device-newer-p state)
    (unless *cuda-lazy-mirror*
      (error "The amplitudes live on the GPU and *CUDA-LAZY-MIRROR* is NIL."))
    (let ((mirror (%mirror state)))
      (cffi:with-pointer-to-vector-data (p mirror)
        (download (device-handle state) p 0 (length mirror))))
    (setf (device-newer-p state) nil)))

(defun sync-to-device (state)
  (when (host-newer-p state)
    (let ((mirror (%mirror state)))
      (cffi:with-pointer-to-vector-data (p mirror)
        (upload (device-handle state) p 0 (length mirror))))
    (setf (host-newer-p state) nil)))

(defmethod qvm::state-elements :before ((state device-state-mixin))
  ;; the host is about to look at (and possibly write) the amplitudes
  (sync-to-host state)
  (setf (host-newer-p state) t))

(defmethod (setf qvm::state-elements) :after (new-value (state device-state-mixin))
  (declare (ignore new-value))
  (setf (host-newer-p state) t
        (device-newer-p state) nil)
  (sync-to-device state))

(defmethod qvm::set-to-zero-state ((state device-state-mixin))
  (setf (gate-tape state) nil)
  (set-zero-state (device-handle state))
  (setf (device-newer-p state) t
        (host-newer-p state) nil))

;;; ---- machines ---------------------------------------------------------------------------------
;;;
;;; The machines are SUBCLASSES of the reference's: every method of the run loop (LOAD-PROGRAM, RUN, TRANSITION, classical
;;; memory; src/execution.lisp, src/transition.lisp) is inherited unchanged, and the protocols whose default methods would
;;; generate host code -- COMPILE-LOADED-PROGRAM / COMPILE-INSTRUCTION -- can be specialised on them (compile.lisp), the way
;;; the reference itself opts DENSITY-QVM out (src/density-qvm.lisp:183-190).

(defclass cuda-pure-state-qvm (qvm:pure-state-qvm) ()
  (:documentation "PURE-STATE-QVM (src/qvm.lisp:114-148) whose STATE is a DEVICE-PURE-STATE."))

(defclass cuda-density-qvm (qvm:density-qvm) ()
  (:documentation "DENSITY-QVM (src/density-qvm.lisp:33-71) whose STATE is a DEVICE-DENSITY-MATRIX-STATE."))

(defvar *cuda-mirror-refresh-limit* 26
  "RUN refreshes the host mirror afterwards only for states of at most this many qubits (callers of the reference's tests
hold the AMPLITUDES vector across runs, tests/gate-tests.lisp:91-103).  Larger states are downloaded lazily, when
QVM::AMPLITUDES / STATE-ELEMENTS is actually called: 16 GiB at 30 qubits should not cross PCIe after every RUN.")

(defun make-cuda-qvm (num-qubits &rest args)
  "QVM:MAKE-QVM (src/qvm.lisp:150-164) with a device-resident state; everything above the state
(LOAD-PROGRAM, RUN, TRANSITION, classical memory) is the unchanged reference code."
  (apply #'make-instance 'cuda-pure-state-qvm
         :number-of-qubits num-qubits
         :state (make-device-pure-state num-qubits)
         args))

(defun make-cuda-density-qvm (num-qubits &rest args)
  "QVM::MAKE-DENSITY-QVM (src/density-qvm.lisp:52-71) with vec(rho) on the device."
  (apply #'make-instance 'cuda-density-qvm
         :number-of-qubits num-qubits
         :state (make-device-density-matrix-state num-qubits)
         args))

(defun %state-index-bits (state)
  (etypecase state
    (device-pure-state (qvm::num-qubits state))
    (device-density-matrix-state (* 2 (qvm::num-qubits state)))))

(defun %before-run (qvm)
  ;; the host may have written into the mirror since it was handed out
  (sync-to-device (qvm::state qvm)))

(defun %after-run (qvm)
  ;; all queued transitions reach the device; the mirror follows only for small states (see *CUDA-MIRROR-REFRESH-LIMIT*)
  (let ((state (qvm::state qvm)))
    (flush-gate-tape state)
    (when (and *cuda-lazy-mirror*
               (<= (%state-index-bits state) *cuda-mirror-refresh-limit*))
      (sync-to-host state))))

;;; Specialised on OUR classes only: an :AFTER method on BASE-QVM would replace the reference's own
;;; (RUN :AFTER BASE-QVM), src/execution.lisp:39-44.
(defmethod qvm:run :before ((qvm cuda-pure-state-qvm)) (%before-run qvm))
(defmethod qvm:run :before ((qvm cuda-density-qvm)) (%before-run qvm))
(defmethod qvm:run :after ((qvm cuda-pure-state-qvm)) (%after-run qvm))
(defmethod qvm:run :after ((qvm cuda-density-qvm)) (%after-run qvm))

;;; REQUIRES-SWAPPING-AMPS-P (src/state-representation.lisp:137-146) compares AMPLITUDES with ORIGINAL-AMPLITUDES; the
;;; stochastic-Kraus path that swaps them is not routed to the device by this shim (SURVEY section 8f #2 is served by the
;;; Python twin), so a device state never needs the swap.
(defmethod qvm::requires-swapping-amps-p ((state device-state-mixin)) nil)
