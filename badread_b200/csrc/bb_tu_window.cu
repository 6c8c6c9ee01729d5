// bb_tu_window.cu — compiles the lane-mode window aligners bb_k_window_lane<4>, <8> (bb_loop.cuh).
#include "bb_launch.h"

cudaError_t bbl_window_lane_init() {
    cudaError_t e = cudaFuncSetAttribute(bb_k_window_lane<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BB_WIN_SMEM_BYTES(4));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(bb_k_window_lane<BB_WIN_LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, BB_WIN_SMEM_BYTES(BB_WIN_LW));
    return e;
}

void bbl_window_lane4(int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, const BBWinTask *tasks, const int *n_tasks,
                      unsigned long long seed, uint32_t *ckpt_pool, uint8_t *tbuf_pool, int *cursor, BBWinTask *fallback,
                      int *fallback_count) {
    bb_k_window_lane<4><<<grid, 64, BB_WIN_SMEM_BYTES(4), st>>>(B, em, tasks, n_tasks, seed, ckpt_pool, tbuf_pool, cursor, fallback,
                                                                fallback_count);
}

void bbl_window_lane8(int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, const BBWinTask *tasks, const int *n_tasks,
                      unsigned long long seed, uint32_t *ckpt_pool, uint8_t *tbuf_pool, int *cursor, BBWinTask *fallback,
                      int *fallback_count) {
    bb_k_window_lane<BB_WIN_LW><<<grid, 64, BB_WIN_SMEM_BYTES(BB_WIN_LW), st>>>(B, em, tasks, n_tasks, seed, ckpt_pool, tbuf_pool, cursor,
                                                                                fallback, fallback_count);
}
