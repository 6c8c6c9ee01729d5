// bb_tu_node_quad.cu — compiles bb_k_node_quad (bb_tasks.cuh): wide-band Hirschberg nodes by a CTA of 8 warps.
#include "bb_launch.h"

cudaError_t bbl_node_quad_init() {
    return cudaFuncSetAttribute(bb_k_node_quad<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BB_QUAD_SMEM_BYTES);
}

void bbl_node_quad(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int parity, int *cursor,
                   int warp_base) {
    bb_k_node_quad<0><<<grid, BB_QUAD_THREADS, BB_QUAD_SMEM_BYTES, st>>>(B, Q, pool, parity, cursor, warp_base);
}
