#include "dm_common.cuh"
#include "dm_internal.h"
bool dm_correlation_umma_supported(int, int) { return false; }
int dm_correlation_umma(const void*, const float*, const void*, const float*, int, int, int, int, float*, cudaStream_t) {
    dm_set_error("tcgen05 engine not built"); return DM_ERR_UNSUPPORTED;
}
