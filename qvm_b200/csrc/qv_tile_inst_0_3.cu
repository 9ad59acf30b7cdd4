// tile kernel, mode 0 (local), 8 amplitudes per thread
#define QV_INST_MODE 0
#define QV_INST_M 3
#include "qv_tile_inst.cuh"
