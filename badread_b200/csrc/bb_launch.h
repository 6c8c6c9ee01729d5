// bb_launch.h — host-side launchers of the heavy kernels.  Every kernel is a template (bb_kernels.cuh), so it is
// compiled only by the translation unit that launches it: the heavy ones live in bb_tu_*.cu, one file each, and
// build in parallel (the wide wavefront instantiations take minutes of ptxas each); bb_api.cu launches the light
// ones itself.  All launchers are asynchronous on `st`; errors surface through cudaGetLastError() in the caller.
#pragma once
#include <cuda_runtime.h>

#include "bb_kernels.cuh"

void bbl_mutate(int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, unsigned long long seed, int *work_counter,
                const int *order, int n_items, bool chain);  // chain: the low-latency build (bb_k_mutate_chain)
cudaError_t bbl_window_lane_init();
void bbl_window_lane4(int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, const BBWinTask *tasks, const int *n_tasks,
                      unsigned long long seed, uint32_t *ckpt_pool, uint8_t *tbuf_pool, int *cursor, BBWinTask *fallback,
                      int *fallback_count);
void bbl_window_lane8(int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, const BBWinTask *tasks, const int *n_tasks,
                      unsigned long long seed, uint32_t *ckpt_pool, uint8_t *tbuf_pool, int *cursor, BBWinTask *fallback,
                      int *fallback_count);
void bbl_window_lane_hist(int words, int ring_t, int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, const BBWinTask *tasks,
                          const int *n_tasks, unsigned long long seed, uint2 *hist_pool, uint8_t *tbuf_pool, int *cursor,
                          BBWinTask *fallback, int *fallback_count);
void bbl_window_warp(int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, BBScratchPool pool, const BBWinTask *tasks,
                     const int *n_tasks, unsigned long long seed, int *cursor);
void bbl_node_warp(int words, int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int parity, int *cursor,
                   int warp_base);  // words per lane: 1, 2 or 4 (class BBQ_NODE_LEAN1 / 2 / 4)
void bbl_node_lane8(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, int parity, int *cursor);
cudaError_t bbl_node_quad_init();
void bbl_node_quad(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int parity, int *cursor,
                   int warp_base);
cudaError_t bbl_node_pair_init();
void bbl_node_pair(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int parity, int *cursor,
                   int warp_base);
void bbl_leaf_warp(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, BBScratchPool pool, int *cursor, int warp_base);
void bbl_leaf_lane_hist(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, uint2 *hist_pool, int *cursor);
cudaError_t bbl_leaf_lane_init();
void bbl_leaf_lane(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, uint32_t *ckpt_pool, int *cursor);
void bbl_align_pair(cudaStream_t st, const uint8_t *q, int n, const uint8_t *t, int m, int k_upper, BBScratchPool pool,
                    uint8_t *ops, unsigned int *dcnt, int *out5);
