// placeholder until the fused tcgen05 solver lands
#include "dm_common.cuh"
#include "dm_internal.h"
bool dm_fused_supported(int, int, int) { return false; }
size_t dm_fused_workspace(char*, int, int, int, int, int, void*) { return 0; }
int dm_fused_solve_chunk(dm_ctx*, const dm_fused_args*, int) { dm_set_error("fused path not built"); return DM_ERR_UNSUPPORTED; }
