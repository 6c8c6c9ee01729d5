// bb_tu_node_pair.cu — compiles bb_k_node_pair (bb_tasks.cuh): wide-band Hirschberg nodes by warp pairs.
#include "bb_launch.h"

cudaError_t bbl_node_pair_init() {
    return cudaFuncSetAttribute(bb_k_node_pair<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BB_PAIR_SMEM_BYTES);
}

void bbl_node_pair(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int parity, int *cursor,
                   int warp_base) {
    bb_k_node_pair<0><<<grid, BB_WARPS_PER_CTA * 32, BB_PAIR_SMEM_BYTES, st>>>(B, Q, pool, parity, cursor, warp_base);
}
