// bb_tu_window_warp.cu — compiles bb_k_window_warp (bb_loop.cuh): windows beyond the lane limits, general aligner.
#include "bb_launch.h"

void bbl_window_warp(int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, BBScratchPool pool, const BBWinTask *tasks,
                     const int *n_tasks, unsigned long long seed, int *cursor) {
    bb_k_window_warp<<<grid, BB_WARPS_PER_CTA * 32, 0, st>>>(B, em, pool, tasks, n_tasks, seed, cursor);
}
