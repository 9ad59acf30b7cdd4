;;;; measurement.lisp -- the measurement protocol on device states (src/measurement.lisp:7-225).
;;;; The random draws stay in Lisp (mt19937, src/measurement.lisp:99,136): the library only ever
;;;; receives uniforms.

(in-package #:qvm-cuda)

(defmethod qvm::get-excited-state-probability ((state device-pure-state) qubit)
  (flush-gate-tape state)
  (call-returning-double #'prob-excited (device-handle state) qubit))

(defmethod qvm::get-excited-state-probability ((state device-density-matrix-state) qubit)
  (flush-gate-tape state)
  (call-returning-double #'density-prob-excited (device-handle state) (qvm::num-qubits state) qubit))

(defmethod qvm::force-measurement (measured-value qubit (state device-pure-state) excited-probability)
  ;; src/measurement.lisp:10-41
  (flush-gate-tape state)
  (let ((inv-norm (if (= 1 measured-value)
                      (/ (sqrt excited-probability))
                      (/ (sqrt (- (qvm:flonum 1) excited-probability))))))
    (collapse (device-handle state) qubit measured-value inv-norm)
    (setf (device-newer-p state) t)
    state))

(defmethod qvm::force-measurement (measured-value qubit (state device-density-matrix-state)
                                   excited-probability)
  ;; src/measurement.lisp:43-68: rescale by 1/p, not 1/sqrt(p)
  (flush-gate-tape state)
  (let ((inv-norm (if (= 1 measured-value)
                      (/ excited-probability)
                      (/ (- (qvm:flonum 1) excited-probability)))))
    (density-collapse (device-handle state) (qvm::num-qubits state) qubit measured-value inv-norm)
    (setf (device-newer-p state) t)
    state))

(defmethod qvm::measure-all-state ((state device-pure-state) (qvm qvm::base-qvm))
  ;; src/measurement.lisp:128-143: one uniform, smallest b with C(b) > p, psi <- |b>
  (flush-gate-tape state)
  (let ((basis-state
          (cffi:with-foreign-objects ((u :double) (out :uint64))
            (setf (cffi:mem-ref u :double) (qvm:flonum (random 1.0d0)))
            (sample (device-handle state) u 1 out 1)
            (cffi:mem-ref out :uint64))))
    (set-basis-state (device-handle state) basis-state)
    (setf (device-newer-p state) t
          (host-newer-p state) nil)
    (values qvm (loop :for i :below (qvm:number-of-qubits qvm)
                      :collect (ldb (byte 1 i) basis-state)))))

(defmethod qvm::apply-measure-discard-to-state (qvm (state device-density-matrix-state)
                                                (instr quil:measure-discard))
  ;; src/measurement.lisp:111-120
  (flush-gate-tape state)
  (density-measure-discard (device-handle state) (qvm::num-qubits state)
                           (quil:qubit-index (quil:measurement-qubit instr)))
  (setf (device-newer-p state) t)
  qvm)

(defun sample-wavefunction-multiple-times/cuda (state num-samples)
  "SAMPLE-WAVEFUNCTION-MULTIPLE-TIMES (src/measurement.lisp:246-288) on a device state: NUM-SAMPLES
uniforms are drawn here, the prefix structure and the searches run on the GPU."
  (flush-gate-tape state)
  (let ((samples (make-array num-samples :element-type '(unsigned-byte 64) :initial-element 0)))
    (cffi:with-foreign-objects ((u :double num-samples) (out :uint64 num-samples))
      (dotimes (i num-samples)
        (setf (cffi:mem-aref u :double i) (random 1.0d0)))
      (sample (device-handle state) u num-samples out 0)
      (dotimes (i num-samples samples)
        (setf (aref samples i) (cffi:mem-aref out :uint64 i))))))

(defun pure-state-expectation/cuda (qvm prepared-state op &optional first-time)
  "PURE-STATE-EXPECTATION (app/src/api/expectation.lisp:78-91) with both wavefunctions on the device:
PREPARED-STATE is a DEVICE-PURE-STATE holding a copy of the prepared wavefunction; OP is run on QVM's own
state and <prepared | OP prepared> is reduced on the GPU instead of the host LOOP :sum."
  (unless first-time
    (copy-state (device-handle (qvm::state qvm)) (device-handle prepared-state))
    (setf (device-newer-p (qvm::state qvm)) t))
  (qvm:load-program qvm op)
  (qvm:run qvm)
  (flush-gate-tape (qvm::state qvm))
  (cffi:with-foreign-object (out :double 2)
    (inner-product (device-handle prepared-state) (device-handle (qvm::state qvm)) out)
    (complex (cffi:mem-aref out :double 0) (cffi:mem-aref out :double 1))))

(defun probabilities/cuda (state)
  "What PERFORM-PROBABILITIES collects (app/src/api/probabilities.lisp): the probability of every basis state of a
DEVICE-PURE-STATE as a vector of double-floats, reduced to 8 bytes per basis state on the device."
  (flush-gate-tape state)
  (let* ((n (expt 2 (qvm::num-qubits state)))
         (probs (make-array n :element-type 'double-float)))
    (cffi:with-pointer-to-vector-data (p probs)
      (probabilities (device-handle state) p 0 n))
    probs))

(defun density-measurement-probabilities/cuda (state)
  "DENSITY-MATRIX-STATE-MEASUREMENT-PROBABILITIES (src/state-representation.lisp:268-286; DENSITY-QVM-MEASUREMENT-PROBABILITIES,
src/density-qvm.lisp:151-154) for a DEVICE-DENSITY-MATRIX-STATE: the real parts of the diagonal of rho, gathered on the device
(2^n doubles cross PCIe instead of the 4^n complex entries of the mirror)."
  (flush-gate-tape state)
  (let* ((dim (expt 2 (qvm::num-qubits state)))
         (probs (make-array dim :element-type 'double-float)))
    (cffi:with-pointer-to-vector-data (p probs)
      (density-diag-probs (device-handle state) (qvm::num-qubits state) p))
    probs))
