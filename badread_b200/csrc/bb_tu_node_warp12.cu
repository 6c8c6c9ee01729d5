// bb_tu_node_warp12.cu — compiles bb_k_node_warp<1> and <2> (bb_tasks.cuh): the narrow single-warp node kernels.
#include "bb_launch.h"

void bbl_node_warp_12(int words, int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int parity,
                      int *cursor, int warp_base) {
    if (words == 2) bb_k_node_warp<2><<<grid, BB_WARPS_PER_CTA * 32, 0, st>>>(B, Q, pool, parity, cursor, warp_base);
    else bb_k_node_warp<1><<<grid, BB_WARPS_PER_CTA * 32, 0, st>>>(B, Q, pool, parity, cursor, warp_base);
}
