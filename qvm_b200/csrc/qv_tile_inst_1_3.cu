// tile kernel, mode 1 (peer), 8 amplitudes per thread
#define QV_INST_MODE 1
#define QV_INST_M 3
#include "qv_tile_inst.cuh"
