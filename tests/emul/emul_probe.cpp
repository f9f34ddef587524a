// TEST INFRASTRUCTURE: lets the tests ask the host build whether the TMA stencil kernels really ran (see cuda_runtime.h)
#include <cuda_runtime.h>

extern "C" long long emul_tma_load_count(void) { return g_emul_tma_loads; }
