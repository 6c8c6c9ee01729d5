// bb_tu_node_warp.cu — compiles bb_k_node_warp<4> (bb_tasks.cuh): Hirschberg nodes by single warps.
#include "bb_launch.h"

void bbl_node_warp4(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int cls, int parity, int *cursor,
                    int warp_base) {
    bb_k_node_warp<4><<<grid, BB_WARPS_PER_CTA * 32, 0, st>>>(B, Q, pool, cls, parity, cursor, warp_base);
}
