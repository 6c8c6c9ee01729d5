// bb_tu_pair.cu — compiles bb_k_align_pair (bb_kernels.cuh): edlib.align for one pair (diagnostics / single-read API).
#include "bb_launch.h"

void bbl_align_pair(cudaStream_t st, const uint8_t *q, int n, const uint8_t *t, int m, int k_upper, BBScratchPool pool,
                    uint8_t *ops, unsigned int *dcnt, int *out5) {
    bb_k_align_pair<<<1, 32, 0, st>>>(q, n, t, m, k_upper, pool, ops, dcnt, out5);
}
