;;;; operators.lisp -- APPLY-GATE-TO-STATE on device states (src/apply-gate.lisp:106-212).
;;;;
;;;; Gates are not sent one by one: every transition appends an entry to the state's tape and the tape is flushed --
;;;; ONE qvmcuda_apply_gates (pure state) or qvmcuda_density_apply_ops (density matrix) call, scheduled, fused and
;;;; compiled inside the library -- when something needs the amplitudes (a measurement, AMPLITUDES, the end of RUN).
;;;; That is the GPU counterpart of COMPILE-LOADED-PROGRAM + fuse-gates (src/qvm.lisp:166-175), and it removes the
;;;; per-transition FFI + launch overhead that dominates 20-qubit programs.
;;;;
;;;; Call stack in INTERPRETED mode (src/transition.lisp:160-180):
;;;;   TRANSITION (pure-state-qvm gate-application) -> (apply #'apply-gate-to-state operator (state qvm) qubits params)
;;;;   -> the (QUIL:GATE DEVICE-PURE-STATE) method below -> %PUSH-GATE.
;;;; Call stack in COMPILED mode (*COMPILE-BEFORE-RUNNING* = T, as the reference's test-suite runs, tests/suite.lisp:13):
;;;;   LOAD-PROGRAM -> COMPILE-LOADED-PROGRAM: for our machines the method in compile.lisp, which leaves the
;;;;   instructions uncompiled, so the interpreted stack above applies.  If compiled instructions reach a device state
;;;;   anyway (a program compiled by another machine, a hand-built PURE-STATE-QVM around a device state),
;;;;   TRANSITION (pure-state-qvm compiled-gate-application) calls (apply-gate-to-state instr (state qvm) NIL)
;;;;   (src/transition.lisp:182-186): QUBITS is NIL there, the reference's own methods ignore it
;;;;   (src/apply-gate.lisp:162-182); the methods below take the qubits from the instruction itself and NEVER call the
;;;;   compiled host lambda (that would run CPU code on the host mirror).

(in-package #:qvm-cuda)

(defun %gate-matrix (gate parameters)
  "Row-major (SIMPLE-ARRAY CFLONUM (d d)) of GATE, as the reference builds it for every transition
(src/apply-gate.lisp:109-114)."
  (qvm::magicl-matrix-to-quantum-operator (apply #'quil:gate-matrix gate parameters)))

(defun %push-gate (state matrix qubits)
  (push (cons matrix (coerce qubits 'list)) (gate-tape state))
  (setf (device-newer-p state) t)
  state)

(defun %instruction-qubits (instr)
  "Qubit indices of a (compiled) gate application in Quil argument order (src/transition.lisp:166)."
  (mapcar #'quil:qubit-index (quil:application-arguments instr)))

(defun %fill-matrix (buffer offset matrix)
  "Copy a row-major complex MATRIX into BUFFER (foreign doubles) at OFFSET as (re, im) pairs; returns the new offset."
  (dotimes (i (array-total-size matrix) offset)
    (let ((z (row-major-aref matrix i)))
      (setf (cffi:mem-aref buffer :double offset) (realpart z)
            (cffi:mem-aref buffer :double (1+ offset)) (imagpart z))
      (incf offset 2))))

(defmethod flush-gate-tape ((state device-pure-state))
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
                      (setf mi (%fill-matrix matrices mi matrix))))
          (apply-gates (device-handle state) n ks qubits matrices
                       (if qvm:*fuse-gates-during-compilation* +fuse+ 0)))))
    state))

;;; Density matrices: every tape entry is ((K_0 K_1 ...) . qubits): a plain gate is the one-element list (SINGLE-KRAUS),
;;; a KRAUS-LIST its operators.  The library turns a whole stretch of them into fused passes over vec(rho): conj(U) on the
;;; column bits, U on the row bits and the superoperator of the channel that follows on the same qubit end up as ONE 4x4
;;; (the reference: 2 passes per unitary and 4m + 3 per m-operator channel, src/apply-gate.lisp:42-99).
(defmethod flush-gate-tape ((state device-density-matrix-state))
  (let ((ops (reverse (gate-tape state))))
    (when ops
      (setf (gate-tape state) nil)
      (sync-to-device state)
      (let* ((n (length ops))
             (total-qubits (reduce #'+ ops :key (lambda (o) (length (cdr o)))))
             (total-doubles (reduce #'+ ops :key (lambda (o)
                                                   (reduce #'+ (car o) :key (lambda (m) (* 2 (array-total-size m))))))))
        (cffi:with-foreign-objects ((ks :int32 n)
                                    (ms :int32 n)
                                    (qubits :int32 total-qubits)
                                    (kraus :double total-doubles))
          (let ((qi 0) (mi 0))
            (loop :for (matrices . qs) :in ops
                  :for i :from 0
                  :do (setf (cffi:mem-aref ks :int32 i) (length qs)
                            (cffi:mem-aref ms :int32 i) (length matrices))
                      (dolist (q (reverse qs))
                        (setf (cffi:mem-aref qubits :int32 qi) q)
                        (incf qi))
                      (dolist (m matrices)
                        (setf mi (%fill-matrix kraus mi m)))))
          (density-apply-ops (device-handle state) (qvm::num-qubits state) n ks qubits ms kraus
                             (if qvm:*fuse-gates-during-compilation* +fuse+ 0)))))
    state))

;;; ---- pure states ------------------------------------------------------------------------------------------------

;;; Every gate class of src/apply-gate.lisp:109-160 funnels into QUIL:GATE-MATRIX.
(defmethod qvm::apply-gate-to-state ((gate quil:gate) (state device-pure-state) qubits &rest parameters)
  (%push-gate state (%gate-matrix gate parameters) qubits))

;;; Compiled gate applications (src/compile-gate.lisp:363-409).  QUBITS is NIL when the caller is TRANSITION
;;; (src/transition.lisp:182-186): the instruction is a QUIL:GATE-APPLICATION and carries its own arguments.
(defmethod qvm::apply-gate-to-state ((gate qvm::compiled-matrix-gate-application)
                                     (state device-pure-state) qubits &rest parameters)
  (declare (ignore parameters))
  (%push-gate state (qvm::compiled-matrix gate) (or qubits (%instruction-qubits gate))))

(defmethod qvm::apply-gate-to-state ((gate qvm::compiled-inlined-matrix-gate-application)
                                     (state device-pure-state) qubits &rest parameters)
  ;; the inlined variant keeps its matrix too (slot GATE-MATRIX, reader COMPILED-MATRIX, src/compile-gate.lisp:397-402);
  ;; its APPLY-OPERATOR is a host lambda and is not called
  (declare (ignore parameters))
  (%push-gate state (qvm::compiled-matrix gate) (or qubits (%instruction-qubits gate))))

(defmethod qvm::apply-gate-to-state ((gate qvm::compiled-permutation-gate-application)
                                     (state device-pure-state) qubits &rest parameters)
  ;; no matrix slot: rebuild it from the source gate, as the interpreted path does for QUIL:PERMUTATION-GATE
  ;; (src/apply-gate.lisp:134-139).  The library recognises 0/1 matrices and runs them as data movement.
  (declare (ignore parameters))
  (%push-gate state (%gate-matrix (qvm::source-gate gate) nil) (or qubits (%instruction-qubits gate))))

;;; ---- density matrices -------------------------------------------------------------------------------------------
;;; gate -> SINGLE-KRAUS, KRAUS-LIST -> one superoperator (src/apply-gate.lisp:42-99,196-212)
(defun %kraus-matrices (sop parameters)
  (adt:match qvm::superoperator sop
    ((qvm::single-kraus u) (list (%gate-matrix u parameters)))
    ((qvm::kraus-list list) (loop :for k :in list :append (%kraus-matrices k parameters)))))

(defmethod qvm::apply-gate-to-state ((gate qvm::superoperator) (state device-density-matrix-state)
                                     qubits &rest parameters)
  (push (cons (%kraus-matrices gate parameters) (coerce qubits 'list)) (gate-tape state))
  (setf (device-newer-p state) t)
  state)

(defmethod qvm::apply-gate-to-state ((gate quil:gate) (state device-density-matrix-state)
                                     qubits &rest parameters)
  (push (cons (list (%gate-matrix gate parameters)) (coerce qubits 'list)) (gate-tape state))
  (setf (device-newer-p state) t)
  state)
