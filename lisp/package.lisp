;;;; package.lisp

(defpackage #:qvm-cuda
  (:use #:cl)
  (:export #:cuda-allocation
           #:device-pure-state
           #:device-density-matrix-state
           #:make-device-pure-state
           #:make-device-density-matrix-state
           #:make-cuda-qvm
           #:make-cuda-density-qvm
           #:cuda-pure-state-qvm
           #:cuda-density-qvm
           #:*cuda-mirror-refresh-limit*
           #:*cuda-device*
           #:*cuda-lazy-mirror*
           #:flush-gate-tape
           #:sample-wavefunction-multiple-times/cuda
           #:pure-state-expectation/cuda
           #:probabilities/cuda
           #:density-measurement-probabilities/cuda))
