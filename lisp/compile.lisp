;;;; compile.lisp -- the compile protocol on machines with device states (src/compile-gate.lisp:409-526,
;;;; src/qvm.lisp:166-175).
;;;;
;;;; The reference turns every gate of a loaded program into a host function (COMPILE-LAMBDA of generated Lisp code) and
;;;; MEASURE into a COMPILED-MEASUREMENT whose projector runs on (AMPLITUDES QVM).  Neither may happen for a device state:
;;;; both would execute CPU code on the host mirror.  The device-side equivalent of that compilation lives in libqvmcuda --
;;;; the scheduler packs the queued gates into fused passes and the pass compiler turns each pass into its own kernel,
;;;; cached by structure -- and is triggered when the tape is flushed.  So, like DENSITY-QVM (src/density-qvm.lisp:183-190),
;;;; our machines keep their programs uncompiled on the Lisp side.

(in-package #:qvm-cuda)

(defmethod qvm::compile-loaded-program ((qvm cuda-pure-state-qvm))
  ;; No QUIL::FUSE-GATES-IN-EXECUTABLE-CODE either: fusion happens in the library (QVMCUDA_FUSE), on the whole run of gates
  ;; between two measurements.
  (setf (qvm::program-compiled-p qvm) t)
  qvm)

(defmethod qvm::compile-instruction ((qvm cuda-pure-state-qvm) isn)
  (declare (ignore qvm))
  isn)

;;; CUDA-DENSITY-QVM inherits DENSITY-QVM's opt-out methods (src/density-qvm.lisp:183-190).

;;; ---- compiled MEASURE that reaches a device state anyway ---------------------------------------------------------
;;; TRANSITION (pure-state-qvm compiled-measurement) funcalls the projector on (AMPLITUDES QVM)
;;; (src/transition.lisp:188-194).  For a device state the same decision rule runs on the device instead
;;; (src/compile-gate.lisp:221-254: p0 = ground-state probability, keep |0> iff (random 1) < p0, rescale by
;;; 1/sqrt(p0) or 1/sqrt(1-p0)); the uniform is drawn here, by the host's generator.
(defun %compiled-measure-on-device (state qubit)
  (flush-gate-tape state)
  (let* ((p0 (call-returning-double #'prob-ground (device-handle state) qubit))
         (keep-zero (< (random (qvm:flonum 1)) p0))
         (bit (if keep-zero 0 1))
         (inv-norm (if keep-zero
                       (/ (sqrt p0))
                       (/ (sqrt (- (qvm:flonum 1) p0))))))
    (collapse (device-handle state) qubit bit inv-norm)
    (setf (device-newer-p state) t)
    bit))

(defmethod qvm:transition :around ((qvm qvm:pure-state-qvm) (instr qvm::compiled-measurement))
  (let ((state (qvm::state qvm)))
    (cond
      ((typep state 'device-pure-state)
       (let ((bit (%compiled-measure-on-device state (quil:qubit-index (quil:measurement-qubit instr))))
             (src (qvm::source-instruction instr)))
         (when (typep src 'quil:measure)
           (setf (qvm::dereference-mref qvm (quil:measure-address src)) bit)))
       (incf (qvm::pc qvm))
       qvm)
      (t (call-next-method)))))
