// bb_tu_mutate.cu — compiles bb_k_mutate and bb_k_mutate_chain (bb_loop.cuh).
#include "bb_launch.h"

// chain: the latency build for the head batch of the longest reads
void bbl_mutate(int grid, cudaStream_t st, BBBatchDev B, BBErrorModelDev em, unsigned long long seed, int *work_counter,
                const int *order, int n_items, bool chain) {
    if (chain) bb_k_mutate_chain<<<grid, BB_WARPS_PER_CTA * 32, 0, st>>>(B, em, seed, work_counter, order, n_items);
    else bb_k_mutate<<<grid, BB_MUTP_THREADS, 0, st>>>(B, em, seed, work_counter, order, n_items);
}
