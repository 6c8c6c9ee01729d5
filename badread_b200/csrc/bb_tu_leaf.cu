// bb_tu_leaf.cu — compiles the leaf kernels bb_k_leaf_warp, bb_k_leaf_lane and bb_k_leaf_lane_hist (bb_tasks.cuh).
#include "bb_launch.h"

void bbl_leaf_warp(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int *cursor, int warp_base) {
    bb_k_leaf_warp<<<grid, BB_WARPS_PER_CTA * 32, 0, st>>>(B, Q, pool, cursor, warp_base);
}

cudaError_t bbl_leaf_lane_init() {
    return cudaFuncSetAttribute(bb_k_leaf_lane<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BB_LEAF_SMEM_BYTES);
}

void bbl_leaf_lane(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, uint32_t *ckpt_pool, int *cursor) {
    bb_k_leaf_lane<0><<<grid, 64, BB_LEAF_SMEM_BYTES, st>>>(B, Q, ckpt_pool, cursor);
}

void bbl_leaf_lane_hist(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, uint2 *hist_pool, int *cursor) {
    bb_k_leaf_lane_hist<0><<<grid, 64, BB_LEAF_RING_BYTES, st>>>(B, Q, hist_pool, cursor);
}
