;;;; qvm-cuda.asd -- B200 back end for the QVM hot path, loaded on top of the unchanged :qvm system.
;;;;
;;;; NOTE: written against quil-lang/qvm v1.18.0; it could not be compiled in the build container
;;;; (no SBCL / quicklisp there).  The same C ABI, called in the same order, is exercised by
;;;; qvm_b200/qvm.py and the tests in tests/.  See INTEGRATION.md.

(asdf:defsystem #:qvm-cuda
  :description "CUDA (sm_100a) engine behind the QVM allocator / state / apply-gate / measurement protocols."
  :license "Apache License 2.0"
  :depends-on (#:qvm #:cffi #:trivial-garbage #:alexandria #:cl-quil)
  :serial t
  :components ((:file "package")
               (:file "bindings")
               (:file "device-state")
               (:file "operators")
               (:file "compile")
               (:file "measurement")))
