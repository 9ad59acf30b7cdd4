;;;; operators.lisp -- APPLY-GATE-TO-STATE on device states (src/apply-gate.lisp:106-212).
;;;;
;;;; Gates are not sent one by one: every transition appends (matrix . qubits) to the state's tape and
;;;; the tape is flushed -- ONE qvmcuda_apply_gates call, scheduled and fused inside the library -- when
;;;; something needs the amplitudes (a measurement, AMPLITUDES, the end of RUN).  That is the GPU
;;;; counterpart of COMPILE-LOADED-PROGRAM + fuse-gates (src/qvm.lisp:166-175), and it removes the
;;;; per-transition FFI + launch overhead that dominates 20-qubit programs.

(in-package #:qvm-cuda)

(defun %gate-matrix (gate parameters)
  "Row-major (SIMPLE-ARRAY CFLONUM (d d)) of GATE, as the reference builds it for every transition
(src/apply-gate.lisp:109-114)."
  (qvm::magicl-matrix-to-quantum-operator (apply #'quil:gate-matrix gate parameters)))

(defun %push-gate (state matrix qubits)
  (push (cons matrix (coerce qubits 'list)) (gate-tape state))
  (setf (device-newer-p state) t)
  state)

(defun flush-gate-tape (state)
  "Send the pending gates to the GPU in one call.  Qubit lists go out in NAT-TUPLE order
(src/utilities.lisp:43-51), i.e. reversed Quil argument order."
  (let ((gates (reverse (gate-tape state))))
    (when gates
      (setf (gate-tape state) nil)
      (sync-to-device state)
      (let* ((n (length gates))
             (total-qubits (reduce #'+ gates :key (lambda (g) (length (cdr g)))))
             (total-doubles (reduce #'+ gates :key (lambda (g) (* 2 (array-total-size (car g)))))))
        (cffi:with-foreign-objects ((ks :int32 n)
                                    (qubits :int32 total-qubits)
                                    (matrices :double total-doubles))
          (let ((qi 0) (mi 0))
            (loop :for (matrix . qs) :in gates
                  :for g :from 0
                  :do (setf (cffi:mem-aref ks :int32 g) (length qs))
                      (dolist (q (reverse qs))
                        (setf (cffi:mem-aref qubits :int32 qi) q)
                        (incf qi))
                      (dotimes (i (array-total-size matrix))
                        (let ((z (row-major-aref matrix i)))
                          (setf (cffi:mem-aref matrices :double mi) (realpart z)
                                (cffi:mem-aref matrices :double (1+ mi)) (imagpart z))
                          (incf mi 2)))))
          (apply-gates (device-handle state) n ks qubits matrices
                       (if qvm:*fuse-gates-during-compilation* +fuse+ 0)))))
    state))

;;; Every gate class of src/apply-gate.lisp:109-160 funnels into QUIL:GATE-MATRIX.
(defmethod qvm::apply-gate-to-state ((gate quil:gate) (state device-pure-state) qubits &rest parameters)
  (%push-gate state (%gate-matrix gate parameters) qubits))

;;; Compiled gate applications carry their matrix (src/compile-gate.lisp:363-467).
(defmethod qvm::apply-gate-to-state ((gate qvm::compiled-matrix-gate-application)
                                     (state device-pure-state) qubits &rest parameters)
  (declare (ignore parameters))
  (%push-gate state (qvm::compiled-matrix gate) qubits))

;;; Density matrices: gate -> SINGLE-KRAUS, KRAUS-LIST -> one superoperator pass
;;; (src/apply-gate.lisp:42-99,196-212).
(defun %kraus-matrices (sop parameters)
  (adt:match qvm::superoperator sop
    ((qvm::single-kraus u) (list (%gate-matrix u parameters)))
    ((qvm::kraus-list list) (loop :for k :in list :append (%kraus-matrices k parameters)))))

(defmethod qvm::apply-gate-to-state ((gate qvm::superoperator) (state device-density-matrix-state)
                                     qubits &rest parameters)
  (sync-to-device state)
  (let* ((kraus (%kraus-matrices gate parameters))
         (k (length qubits))
         (d (expt 2 k))
         (m (length kraus)))
    (cffi:with-foreign-objects ((qs :int32 k)
                                (buf :double (* 2 d d m)))
      (loop :for q :in (reverse (coerce qubits 'list))
            :for i :from 0
            :do (setf (cffi:mem-aref qs :int32 i) q))
      (let ((mi 0))
        (dolist (mat kraus)
          (dotimes (i (* d d))
            (let ((z (row-major-aref mat i)))
              (setf (cffi:mem-aref buf :double mi) (realpart z)
                    (cffi:mem-aref buf :double (1+ mi)) (imagpart z))
              (incf mi 2)))))
      (density-apply-kraus (device-handle state) (qvm::num-qubits state) k qs m buf +fuse+))
    (setf (device-newer-p state) t)
    state))

(defmethod qvm::apply-gate-to-state ((gate quil:gate) (state device-density-matrix-state)
                                     qubits &rest parameters)
  (apply #'qvm::apply-gate-to-state (qvm::single-kraus gate) state qubits parameters))
