// bb_tu_leaf.cu — compiles the leaf kernels bb_k_leaf_warp and bb_k_leaf_lane (bb_tasks.cuh).
#include "bb_launch.h"

void bbl_leaf_warp(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int *cursor, int warp_base) {
    bb_k_leaf_warp<<<grid, BB_WARPS_PER_CTA * 32, 0, st>>>(B, Q, pool, cursor, warp_base);
}

void bbl_leaf_lane(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, uint2 *hist_pool, int *cursor) {
    bb_k_leaf_lane<<<grid, 64, 0, st>>>(B, Q, hist_pool, cursor);
}
