// emu_align.cpp — compiles the device aligner (badread_b200/csrc/bb_align.cuh) for the host through the warp
// emulator and exposes it to the CPU-only tests (TEST INFRASTRUCTURE).
#include "cuda_emu.h"

#include <vector>

#include "../../badread_b200/csrc/bb_align.cuh"

extern "C" __attribute__((visibility("default")))
int emu_align_path(const uint8_t *q, int n, const uint8_t *t, int m, int k_upper, int qabs_pad, int maxl, uint8_t *ops,
                   uint16_t *dcnt, int *out5) {
    // qabs_pad > 0: the query is embedded at offset qabs_pad of a longer "read" (exercises the bitmap offsets)
    std::vector<uint8_t> read((size_t)qabs_pad + n + 64, 'C');
    std::memcpy(read.data() + qabs_pad, q, (size_t)n);
    const int read_len = qabs_pad + n + 7;
    std::vector<uint2> hist(106496);
    std::vector<int8_t> hbuf((size_t)std::max(n, m) + 64);
    std::vector<int> LR(2 * ((size_t)std::max(n, m) + 64)), stack(5 * 64);
    std::vector<uint4> peq((size_t)read_len / 32 + 8);
    BBScratch sc;
    sc.hist = hist.data(); sc.hist_cap = (int)hist.size();
    sc.hbuf = hbuf.data(); sc.hbuf_cap = (int)hbuf.size();
    sc.L = LR.data(); sc.R = LR.data() + LR.size() / 2; sc.lr_cap = (int)(LR.size() / 2);
    sc.stack = stack.data(); sc.stack_cap = 64;
    sc.peq = peq.data(); sc.peq_cap = (int)peq.size();
    int lead = 0;
    BBAlnCounts result = {0, 0, 0, 0};
    std::memset(dcnt, 0, (size_t)n * sizeof(uint16_t));
    emu::run_warp([&]() {
        bb_build_peq(read.data(), read_len, sc.peq);
        BBEmit em = {ops, dcnt, &lead};
        BBAlnCounts cnt = {0, 0, 0, 0};
        const uint8_t *qq = read.data() + qabs_pad;
        if (maxl == 1) bb_align<true, 1>(qq, n, t, m, k_upper, sc, em, qabs_pad, cnt);
        else if (maxl == 2) bb_align<true, 2>(qq, n, t, m, k_upper, sc, em, qabs_pad, cnt);
        else bb_align<true, 16>(qq, n, t, m, k_upper, sc, em, qabs_pad, cnt);
        if ((threadIdx.x & 31) == 0) result = cnt;
    });
    out5[0] = result.matches; out5[1] = result.dels; out5[2] = result.dist; out5[3] = lead; out5[4] = result.err;
    return 0;
}
