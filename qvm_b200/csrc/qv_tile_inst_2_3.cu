// tile kernel, mode 2 (pull), 8 amplitudes per thread
#define QV_INST_MODE 2
#define QV_INST_M 3
#include "qv_tile_inst.cuh"
