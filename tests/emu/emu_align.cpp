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

#include "../../badread_b200/csrc/bb_lane.cuh"

// Lane-mode aligner (one problem per thread): every emulated lane solves the same problem, lane 0 reports.
extern "C" __attribute__((visibility("default")))
int emu_lane_align(const uint8_t *q, int n, const uint8_t *t, int m, int k_upper, int qabs_pad, int lw, int *out4) {
    std::vector<uint8_t> read((size_t)qabs_pad + n + 64, 'G');
    std::memcpy(read.data() + qabs_pad, q, (size_t)n);
    const int read_len = qabs_pad + n + 9;
    std::vector<uint4> peq((size_t)read_len / 32 + 8);
    int a, b;
    {
        const int diff = n > m ? n - m : m - n;
        if (k_upper < diff) k_upper = diff;
        const int mx = n > m ? n : m;
        if (k_upper > mx) k_upper = mx;
    }
    bb_band(n, m, k_upper, a, b);
    if (bb_lane_words(a, b) > lw) return -1;
    std::vector<uint2> hist((size_t)m * lw + 8);
    int res[4] = {0, 0, 0, 0};
    emu::run_warp([&]() {
        bb_build_peq(read.data(), read_len, peq.data());
        if (threadIdx.x != 0) return;
        BBLaneProb P;
        P.peq = peq.data(); P.peq_bit0 = qabs_pad + 32; P.q = read.data() + qabs_pad; P.n = n; P.t = t; P.m = m;
        P.a = a; P.b = b; P.hist = hist.data();
        int d, mt = 0, dl = 0, err = 0;
        if (lw == 4) { d = bb_lane_pass<4>(P); bb_lane_traceback<4>(P, mt, dl, err); }
        else { d = bb_lane_pass<8>(P); bb_lane_traceback<8>(P, mt, dl, err); }
        res[0] = mt; res[1] = dl; res[2] = d; res[3] = err;
    });
    for (int i = 0; i < 4; i++) out4[i] = res[i];
    return 0;
}
