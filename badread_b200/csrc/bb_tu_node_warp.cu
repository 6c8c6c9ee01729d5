// bb_tu_node_warp.cu — compiles bb_k_node_warp<4> (bb_tasks.cuh): Hirschberg nodes by single warps, 4 words per lane.
#include "bb_launch.h"

void bbl_node_warp_12(int words, int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int parity,
                      int *cursor, int warp_base);

void bbl_node_warp(int words, int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int parity, int *cursor,
                   int warp_base) {
    if (words == 4) bb_k_node_warp<4><<<grid, BB_WARPS_PER_CTA * 32, 0, st>>>(B, Q, pool, parity, cursor, warp_base);
    else bbl_node_warp_12(words, grid, st, B, Q, pool, parity, cursor, warp_base);
}
