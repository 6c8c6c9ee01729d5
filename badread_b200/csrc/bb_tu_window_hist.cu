// bb_tu_window_hist.cu — compiles the default lane-mode window aligners bb_k_window_lane_hist<4>, <8> (bb_loop.cuh).
#include "bb_launch.h"

void bbl_window_lane_hist(int words, int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, const BBWinTask *tasks,
                          const int *n_tasks, unsigned long long seed, uint2 *hist_pool, uint8_t *tbuf_pool, int *cursor,
                          BBWinTask *fallback, int *fallback_count) {
    if (words == 4)
        bb_k_window_lane_hist<4><<<grid, 64, 0, st>>>(B, em, tasks, n_tasks, seed, hist_pool, tbuf_pool, cursor, fallback, fallback_count);
    else
        bb_k_window_lane_hist<BB_WIN_LW><<<grid, 64, 0, st>>>(B, em, tasks, n_tasks, seed, hist_pool, tbuf_pool, cursor, fallback,
                                                              fallback_count);
}
