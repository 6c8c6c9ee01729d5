// emu_align.cpp — compiles the device aligner (badread_b200/csrc/bb_align.cuh) for the host through the warp
// emulator and exposes it to the CPU-only tests (TEST INFRASTRUCTURE).
#include "cuda_emu.h"

#include <vector>

#include "../../badread_b200/csrc/bb_kernels.cuh"
#include "../../badread_b200/csrc/bb_models.cuh"

extern "C" __attribute__((visibility("default")))
int emu_align_path(const uint8_t *q, int n, const uint8_t *t, int m, int k_upper, int qabs_pad, int maxl, uint8_t *ops,
                   unsigned int *dcnt, int *out5) {
    // qabs_pad > 0: the query is embedded at offset qabs_pad of a longer "read" (exercises the bitmap offsets)
    std::vector<uint8_t> read((size_t)qabs_pad + n + 64, 'C');
    std::memcpy(read.data() + qabs_pad, q, (size_t)n);
    const int read_len = qabs_pad + n + 7;
    std::vector<uint2> hist(106496);
    std::vector<int8_t> hbuf((size_t)std::max(n, m) + 64);
    std::vector<int> LR(2 * ((size_t)std::max(n, m) + 64)), stack(5 * 64);
    std::vector<uint4> peq((size_t)bb_peq_words(read_len) + 8);
    BBScratch sc;
    sc.hist = hist.data(); sc.hist_cap = (int)hist.size();
    sc.hbuf = hbuf.data(); sc.hbuf_cap = (int)hbuf.size();
    sc.L = LR.data(); sc.R = LR.data() + LR.size() / 2; sc.lr_cap = (int)(LR.size() / 2);
    sc.stack = stack.data(); sc.stack_cap = 64;
    sc.peq = peq.data(); sc.peq_cap = (int)peq.size();
    int lead = 0;
    BBAlnCounts result = {0, 0, 0, 0};
    std::memset(dcnt, 0, (size_t)n * sizeof(unsigned int));
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


// Lane-mode aligner (one problem per thread): every emulated lane solves the same problem, lane 0 reports.
extern "C" __attribute__((visibility("default")))
int emu_lane_align(const uint8_t *q, int n, const uint8_t *t, int m, int k_upper, int qabs_pad, int lw, int *out4) {
    std::vector<uint8_t> read((size_t)qabs_pad + n + 64, 'G');
    std::memcpy(read.data() + qabs_pad, q, (size_t)n);
    const int read_len = qabs_pad + n + 9;
    std::vector<uint4> peq((size_t)bb_peq_words(read_len) + 8);
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
        P.peq = peq.data(); P.peq_bit0 = qabs_pad + BB_PEQ_BIT0; P.q = read.data() + qabs_pad; P.n = n; P.t = t; P.m = m;
        P.a = a; P.b = b; P.hist = hist.data();
        int d, mt = 0, dl = 0, err = 0;
        if (lw == 4) { d = bb_lane_pass<4>(P); bb_lane_traceback<4>(P, mt, dl, err); }
        else { d = bb_lane_pass<8>(P); bb_lane_traceback<8>(P, mt, dl, err); }
        res[0] = mt; res[1] = dl; res[2] = d; res[3] = err;
    });
    for (int i = 0; i < 4; i++) out4[i] = res[i];
    return 0;
}


static int use_quad = 1;  // wide nodes by the 8-warp CTA kernel (1) or by warp pairs (0)
extern "C" __attribute__((visibility("default"))) void emu_set_quad(int v) { use_quad = v; }
static int g_use_hist = 0;  // 1: the default lane builds (global history + shared-memory staging ring) instead of the checkpoint builds
extern "C" __attribute__((visibility("default"))) void emu_set_hist(int v) { g_use_hist = v; }

// The level-synchronous task pipeline (bb_tasks.cuh) for one read, every kernel as one emulated warp.
extern "C" __attribute__((visibility("default")))
int emu_tasks_align(const uint8_t *seq, int n, const uint8_t *frag, int m, int upper, uint8_t *ops, unsigned int *dcnt,
                    int *out5) {
    std::vector<uint8_t> sq(seq, seq + n), fr(frag, frag + m);
    sq.resize((size_t)n + 64, 0); fr.resize((size_t)m + 64, 0);
    std::vector<uint4> speq((size_t)bb_peq_words(n) + 8);
    BBReadDev rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.seq_len = n; rd.frag_len = m; rd.upper = upper;
    BBBatchDev B;
    std::memset(&B, 0, sizeof(B));
    unsigned long long ridx = 0;
    B.n_reads = 1; B.read_index = &ridx; B.reads = &rd; B.frag = fr.data(); B.seq = sq.data(); B.ops = ops; B.dcnt = dcnt;
    B.speq = speq.data();
    std::vector<uint4> fpeq((size_t)bb_peq_words(m) + 8);
    B.fpeq = fpeq.data();  // rd.fpeq_off = 0
    std::memset(dcnt, 0, (size_t)n * sizeof(unsigned int));
    const int cap = 8192;
    std::vector<BBNode> qn[BBQ_NODE_CLASSES][2], ql[2];
    std::vector<int> cnt(512, 0);
    BBQueues Q;
    for (int c = 0; c < BBQ_NODE_CLASSES; c++) for (int p = 0; p < 2; p++) { qn[c][p].resize(cap); Q.node[c][p] = qn[c][p].data(); }
    for (int w = 0; w < 2; w++) { ql[w].resize(cap); Q.leaf[w] = ql[w].data(); }
    Q.count = cnt.data(); Q.overflow = cnt.data() + BBQ_OVERFLOW; Q.cap_node = cap; Q.cap_leaf = cap; Q.lane8_cols = 4096;
    // scratch for the warp kernels (warp 0 only) and the lane leaf kernel
    const int big = std::max(n, m) + 64;
    const int NW = BB_WARPS_PER_CTA;  // scratch for every warp of one emulated CTA
    std::vector<uint2> hist((size_t)NW * 106496), lhist((size_t)32 * BB_LEAF_LANE_COLS * BB_LEAF_LW);
    std::vector<uint32_t> lckpt((size_t)64 * BB_LEAF_MAX_TILES * BB_LEAF_CKPT_WORDS);
    std::vector<int8_t> hbuf((size_t)NW * big);
    std::vector<int> LR((size_t)NW * 2 * big), stack((size_t)NW * 5 * 64);
    std::vector<uint8_t> tbuf(16);
    BBScratchPool pool;
    std::memset(&pool, 0, sizeof(pool));
    pool.hist = hist.data(); pool.hist_stride = 106496; pool.hist_cap = 106496;
    pool.hbuf = hbuf.data(); pool.hbuf_stride = big; pool.hbuf_cap = big;
    pool.lr = LR.data(); pool.lr_stride = 2 * (long long)big; pool.lr_cap = big;
    pool.stack = stack.data(); pool.stack_cap = 64;
    pool.tbuf = tbuf.data(); pool.tbuf_stride = 0;
    pool.peq = speq.data(); pool.peq_stride = 0; pool.peq_cap = (int)speq.size();
    emu::run_warp([&]() { bb_build_peq(sq.data(), n, speq.data()); });
    emu::run_warp([&]() { bb_build_peq(fr.data(), m, fpeq.data()); });
    int order0 = 0;
    emu::run_warp([&]() { bb_k_push_roots(B, Q, Q, &order0); });
    int *cursor = cnt.data() + 16;
    for (int level = 0; level < 40; level++) {
        const int p = level & 1;
        for (int c = 0; c < BBQ_NODE_CLASSES; c++) cnt[BBQ_COUNT(c, p ^ 1)] = 0;
        int *c0 = cursor++, *c1 = cursor++, *c2 = cursor++, *c2b = cursor++, *c2c = cursor++;
        const int po = p | (level > 0 ? BBQ_BACKWARDS : 0);  // as bb_api.cu: queues below the roots are walked from the end
        if (use_quad) emu::run_block(BB_QUAD_THREADS, [&]() { bb_k_node_quad(B, Q, pool, po, c0, 0); });
        else emu::run_block(BB_WARPS_PER_CTA * 32, [&]() { bb_k_node_pair(B, Q, pool, po, c0, 0); });
        emu::run_warp([&]() { bb_k_node_warp<4>(B, Q, pool, po, c1, 0); });
        emu::run_warp([&]() { bb_k_node_warp<2>(B, Q, pool, po, c2, 0); });
        emu::run_warp([&]() { bb_k_node_warp<1>(B, Q, pool, po, c2c, 0); });
        emu::run_warp([&]() { bb_k_node_lane<BB_NODE_LW_SMALL>(B, Q, po, c2b); });
        int pending = 0;
        for (int c = 0; c < BBQ_NODE_CLASSES; c++) pending += cnt[BBQ_COUNT(c, p ^ 1)];
        if (pending == 0) break;
    }
    int *c3 = cursor++, *c4 = cursor++;
    emu::run_warp([&]() { bb_k_leaf_warp(B, Q, pool, c3, 0); });
    if (g_use_hist) emu::run_warp([&]() { bb_k_leaf_lane_hist(B, Q, lhist.data(), c4); });
    else emu::run_warp([&]() { bb_k_leaf_lane(B, Q, lckpt.data(), c4); });
    out5[0] = rd.matches; out5[1] = rd.dels; out5[2] = cnt[BBQ_OVERFLOW]; out5[3] = rd.lead_del; out5[4] = rd.flags;
    return 0;
}


// The two-column bit-plane distance pass (bb_band_pass_bp) against the single-column one on the same problem: both directions side by side in
// two 16-lane groups (K = 16), as the lean warp node kernel runs them.  Returns the number of mismatching outputs
// (column-score entries + the two corner scores), or -1 if the band does not fit L with K = 16.
template <int L>
static int emu_compare_passes_impl(const uint8_t *q, int n, const uint8_t *t, int m, int a, int b) {
    std::vector<uint4> peq((size_t)bb_peq_words(n) + 8), tpeq((size_t)bb_peq_words(m) + 8);
    const int left_w = m / 2, right_w = m - left_w;
    const int loL = std::max(0, left_w - 1 - a), loR = std::max(0, right_w - 1 - a);
    std::vector<int> ref((size_t)2 * (n + 64), -7), got((size_t)2 * (n + 64), -7);
    int corner_ref[2] = {0, 0}, corner_got[2] = {0, 0};
    for (int variant = 0; variant < 2; variant++) {
        std::vector<int> &out = variant ? got : ref;
        int *corner = variant ? corner_got : corner_ref;
        emu::run_warp([&]() {
            const int lane = threadIdx.x & 31;
            bb_build_peq(q, n, peq.data());
            bb_build_peq(t, m, tpeq.data());
            const bool rev = lane >= 16;
            BBProb P;
            P.n = n; P.a = a; P.b = b; P.peq = peq.data(); P.hist = nullptr; P.nb_alloc = 0;
            P.tpeq = tpeq.data();
            if (!rev) {
                P.q = q; P.qs = 1; P.t = t; P.ts = 1; P.ncols = left_w; P.peq_bit0 = BB_PEQ_BIT0;
                P.cols_out = out.data(); P.cols_lo = loL; P.tpeq_bit0 = BB_PEQ_BIT0;
            } else {
                P.q = q + n - 1; P.qs = -1; P.t = t + m - 1; P.ts = -1; P.ncols = right_w; P.peq_bit0 = n - 1 + BB_PEQ_BIT0;
                P.cols_out = out.data() + n + 64; P.cols_lo = loR; P.tpeq_bit0 = m - 1 + BB_PEQ_BIT0;
            }
            const int r = variant ? bb_band_pass_bp<L>(P, 16) : bb_band_pass<L, false, true>(P, 16);
            if (lane == 0) corner[0] = r;
            if (lane == 16) corner[1] = r;
        });
    }
    int bad = 0;
    for (size_t i = 0; i < ref.size(); i++) bad += ref[i] != got[i];
    bad += corner_ref[0] != corner_got[0];
    bad += corner_ref[1] != corner_got[1];
    return bad;
}

extern "C" __attribute__((visibility("default")))
int emu_compare_passes(const uint8_t *q, int n, const uint8_t *t, int m, int k) {
    int a, b;
    bb_band(n, m, k, a, b);
    int bad = 0, done = 0;
    for (int even = 0; even < 2; even++) {   // the band as it is (odd bands: chunks handled column by column at their ends) and
        if (even) { a += a & 1; b += b & 1; }  // made even as bb_task_band does (entry / exit on the common path)
        const int L = bb_pick_L<4>(a, b, 16);
        if (L <= 0) continue;
        done++;
        if (L == 4) bad += emu_compare_passes_impl<4>(q, n, t, m, a, b);
        else if (L == 2) bad += emu_compare_passes_impl<2>(q, n, t, m, a, b);
        else bad += emu_compare_passes_impl<1>(q, n, t, m, a, b);
    }
    return done ? bad : -1;
}


// The shared-memory match-word cache of the wide wavefront (bb_band_pass<L, ., ., true>) against the variant that
// streams the words from the bitmap, forward and reverse pass of one node, one full warp each (K = 32).
template <int L>
static int emu_compare_sm_impl(const uint8_t *q, int n, const uint8_t *t, int m, int a, int b) {
    std::vector<uint4> peq((size_t)bb_peq_words(n) + 8);
    std::vector<uint32_t> esm((size_t)bb_esm_words(L) + 4);
    const int left_w = m / 2, right_w = m - left_w;
    const int loL = std::max(0, left_w - 1 - a), loR = std::max(0, right_w - 1 - a);
    int bad = 0;
    for (int rev = 0; rev < 2; rev++) {
        std::vector<int> out[2] = {std::vector<int>((size_t)n + 64, -7), std::vector<int>((size_t)n + 64, -7)};
        int corner[2] = {0, 0};
        for (int variant = 0; variant < 2; variant++) {
            emu::run_warp([&]() {
                bb_build_peq(q, n, peq.data());
                BBProb P;
                P.n = n; P.a = a; P.b = b; P.peq = peq.data(); P.hist = nullptr; P.nb_alloc = 0;
                P.cols_out = out[variant].data();
                if (!rev) {
                    P.q = q; P.qs = 1; P.t = t; P.ts = 1; P.ncols = left_w; P.peq_bit0 = BB_PEQ_BIT0; P.cols_lo = loL;
                } else {
                    P.q = q + n - 1; P.qs = -1; P.t = t + m - 1; P.ts = -1; P.ncols = right_w;
                    P.peq_bit0 = n - 1 + BB_PEQ_BIT0; P.cols_lo = loR;
                }
                P.esm = esm.data();
                const int r = variant ? bb_band_pass<L, false, true, true>(P, 32) : bb_band_pass<L, false, true, false>(P, 32);
                if ((threadIdx.x & 31) == 0) corner[variant] = r;
            });
        }
        for (size_t i = 0; i < out[0].size(); i++) bad += out[0][i] != out[1][i];
        bad += corner[0] != corner[1];
    }
    return bad;
}

extern "C" __attribute__((visibility("default")))
int emu_compare_sm(const uint8_t *q, int n, const uint8_t *t, int m, int k, int L) {
    int a, b;
    bb_band(n, m, k, a, b);
    if ((a + b) / (32 * L) + 2 > 32) return -1;
    if (L == 8) return emu_compare_sm_impl<8>(q, n, t, m, a, b);
    if (L == 16) return emu_compare_sm_impl<16>(q, n, t, m, a, b);
    if (L == 32) return emu_compare_sm_impl<32>(q, n, t, m, a, b);
    return -1;
}


// The lane-mode window aligner kernel (bb_loop.cuh) on one read: `frag` is the padded fragment, change i (1-based
// ordinal) rewrites slot pos[i] with the inline-encoded string enc[i] (len | c0 << 8 | c1 << 16 | c2 << 24, len <= 3).
// Runs identity re-measurements a = 1 .. n_changes / 25 through bb_k_window_lane<LW> as one emulated warp and returns
// (matches, columns) per measurement; windows that do not fit LW words come back as (-1, -1).
extern "C" __attribute__((visibility("default")))
int emu_window_lane(const uint8_t *frag, int frag_len, const int *pos, const uint32_t *enc, int n_changes,
                    unsigned long long seed, unsigned long long read_index, int lw, int *out_pairs) {
    std::vector<uint8_t> fr(frag, frag + frag_len);
    fr.resize((size_t)frag_len + 64, 0);
    std::vector<uint32_t> state((size_t)frag_len + 64, BB_SLOT_NONE);
    std::vector<unsigned int> ctime((size_t)frag_len + 64, 0u);
    for (int i = 0; i < n_changes; i++) { state[(size_t)pos[i]] = enc[i]; ctime[(size_t)pos[i]] = (unsigned int)(i + 1); }
    std::vector<uint4> fpeq((size_t)bb_peq_words(frag_len) + 8);
    const int n_meas = n_changes / BB_ALIGNMENT_INTERVAL;
    std::vector<int2> wres((size_t)n_meas + 4, make_int2(-1, -1));
    BBReadDev rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.frag_len = frag_len;
    BBBatchDev B;
    std::memset(&B, 0, sizeof(B));
    B.n_reads = 1; B.read_index = &read_index; B.reads = &rd; B.frag = fr.data(); B.state = state.data();
    B.ctime = ctime.data(); B.fpeq = fpeq.data(); B.wres = wres.data();
    BBErrorModelDev em;
    std::memset(&em, 0, sizeof(em));
    em.k = 7; em.type = 1;
    std::vector<BBWinTask> tasks, fallback((size_t)n_meas + 4);
    for (int a = 1; a <= n_meas; a++) tasks.push_back(BBWinTask{0, a});
    int n_tasks = n_meas, cursor = 0, fb_count = 0;
    std::vector<uint32_t> ckpt((size_t)64 * BB_WIN_MAX_TILES * BB_WIN_CKPT_WORDS(BB_WIN_LW));
    std::vector<uint8_t> tbuf((size_t)64 * BB_WIN_MAX_COLS);
    emu::run_warp([&]() { bb_build_peq(fr.data(), frag_len, fpeq.data()); });
    std::vector<uint2> whist(g_use_hist ? (size_t)32 * BB_WIN_MAX_COLS * BB_WIN_LW : 1);
    emu::run_warp([&]() {
        if (g_use_hist == 2 && lw == 4)  // two columns per tick: the ring wraps every four columns
            bb_k_window_lane_hist<4, 2>(B, em, tasks.data(), &n_tasks, seed, whist.data(), tbuf.data(), &cursor, fallback.data(), &fb_count);
        else if (g_use_hist && lw == 4)
            bb_k_window_lane_hist<4, 4>(B, em, tasks.data(), &n_tasks, seed, whist.data(), tbuf.data(), &cursor, fallback.data(), &fb_count);
        else if (g_use_hist)
            bb_k_window_lane_hist<BB_WIN_LW, 4>(B, em, tasks.data(), &n_tasks, seed, whist.data(), tbuf.data(), &cursor, fallback.data(), &fb_count);
        else if (lw == 4) bb_k_window_lane<4>(B, em, tasks.data(), &n_tasks, seed, ckpt.data(), tbuf.data(), &cursor, fallback.data(), &fb_count);
        else bb_k_window_lane<BB_WIN_LW>(B, em, tasks.data(), &n_tasks, seed, ckpt.data(), tbuf.data(), &cursor, fallback.data(), &fb_count);
    });
    for (int a = 0; a < n_meas; a++) { out_pairs[2 * a] = wres[(size_t)a].x; out_pairs[2 * a + 1] = wres[(size_t)a].y; }
    return rd.flags;
}


// The error loop (bb_loop.cuh) of ONE read under the emulator, as bb_api.cu enqueues it: bb_k_mutate -> bb_k_window_tasks ->
// bb_k_window_lane_hist<4> -> <8> -> bb_k_window_warp -> bb_k_replay, round after round until the read is done.  The
// fragment comes unpadded; the pads are drawn like bb_k_build_fragments draws them.  The error model comes as the flat
// tables bb_upload_error_model takes.  out8: loop_count, change_count, n_align, seq_len, start_trim, end_trim, upper, flags;
// joined_out: ''.join(new_fragment_bases) after the loop (cut to joined_cap), checked against bb_k_join's output (read,
// match bitmap, distance bound); padded_out (optional): the fragment with its pads.  Returns the number of rounds, or < 0.
extern "C" __attribute__((visibility("default")))
int emu_error_loop(const uint8_t *fragment, int n, double target, unsigned long long seed, unsigned long long read_index,
                   int k, const int32_t *kmer_to_row, int32_t n_rows, const int32_t *row_off, const double *cum,
                   const uint8_t *flags, const uint32_t *slots, const uint8_t *pool_bytes, int *out8, uint8_t *joined_out,
                   int joined_cap, uint8_t *padded_out) {
    const int frag_len = n + 2 * k;
    std::vector<BBRowInfo> info((size_t)n_rows);
    for (int32_t r = 0; r < n_rows; r++) {
        const int32_t e0 = row_off[r], ne = row_off[r + 1] - e0;
        BBRowInfo &ri = info[(size_t)r];
        ri.cum_last = cum[e0 + ne - 1]; ri.cum0 = cum[e0]; ri.e0 = e0; ri.ne = ne;
        ri.first_is_identity = flags[e0] == 1 ? 1 : 0; ri.pad = 0;
    }
    BBErrorModelDev em;
    std::memset(&em, 0, sizeof(em));
    em.k = k; em.type = 1; em.kmer_to_row = kmer_to_row; em.row_off = row_off; em.cum = cum; em.flags = flags; em.slots = slots;
    em.pool = pool_bytes; em.rowinfo = info.data();
    std::vector<uint8_t> fr((size_t)frag_len + 64, 0);
    {
        BBRng rng;
        rng.init(seed, read_index);
        rng.stream(BB_PURPOSE_PAD, 0);
        for (int j = 0; j < k; j++) fr[(size_t)j] = rng.random_base();
        for (int j = 0; j < k; j++) fr[(size_t)(frag_len - k + j)] = rng.random_base();
        std::memcpy(fr.data() + k, fragment, (size_t)n);
    }
    std::vector<uint32_t> state((size_t)frag_len + 64, BB_SLOT_NONE);
    std::vector<unsigned int> ctime((size_t)frag_len + 64, 0u);
    std::vector<int> kidx((size_t)frag_len + 64, -1);
    for (int x = 0; x + k <= frag_len; x++) {
        int idx = 0;
        bool ok = true;
        for (int j = 0; j < k; j++) {
            const uint8_t c = fr[(size_t)(x + j)];
            const int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
            if (code < 0) ok = false;
            idx = idx * 4 + (code & 3);
        }
        kidx[(size_t)x] = ok ? kmer_to_row[idx] : -1;
    }
    std::vector<uint4> fpeq((size_t)bb_peq_words(frag_len) + 8);
    const int cap = (int)(0.9 * (double)frag_len) + k + 2;
    std::vector<uint2> chlog((size_t)cap + 8);
    std::vector<int2> wres((size_t)cap / BB_ALIGNMENT_INTERVAL + 8, make_int2(-1, -1));
    BBReadDev rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.frag_len = frag_len;
    const double need = (double)frag_len * (1.0 - target);
    rd.horizon = (int)std::min<double>((double)cap, std::max(0.0, 1.25 * need) + 48.0);
    rd.status = BB_READ_PENDING;
    BBBatchDev B;
    std::memset(&B, 0, sizeof(B));
    int order0 = 0;
    B.n_reads = 1; B.read_index = &read_index; B.target = &target; B.order = &order0; B.reads = &rd; B.frag = fr.data();
    B.state = state.data(); B.ctime = ctime.data(); B.kidx = kidx.data(); B.fpeq = fpeq.data(); B.chlog = chlog.data();
    B.wres = wres.data();
    emu::run_warp([&]() { bb_build_peq(fr.data(), frag_len, fpeq.data()); });
    // scratch of the warp window kernel (warp 0 of one CTA) and of the lane kernels (one warp)
    const int big = 2 * BB_WIN_MAX_COLS + frag_len + 64;
    const int NW = BB_WARPS_PER_CTA;
    std::vector<uint2> hist((size_t)NW * 106496), whist((size_t)32 * BB_WIN_MAX_COLS * BB_WIN_LW);
    std::vector<int8_t> hbuf((size_t)NW * big);
    std::vector<int> LR((size_t)NW * 2 * big), stack((size_t)NW * 5 * 64);
    std::vector<uint8_t> wtbuf((size_t)NW * big), ltbuf((size_t)64 * BB_WIN_MAX_COLS);
    std::vector<uint4> wpeq((size_t)NW * (bb_peq_words(big) + 8));
    BBScratchPool pool;
    std::memset(&pool, 0, sizeof(pool));
    pool.hist = hist.data(); pool.hist_stride = 106496; pool.hist_cap = 106496;
    pool.hbuf = hbuf.data(); pool.hbuf_stride = big; pool.hbuf_cap = big;
    pool.lr = LR.data(); pool.lr_stride = 2 * (long long)big; pool.lr_cap = big;
    pool.stack = stack.data(); pool.stack_cap = 64;
    pool.tbuf = wtbuf.data(); pool.tbuf_stride = big;
    pool.peq = wpeq.data(); pool.peq_stride = bb_peq_words(big) + 8; pool.peq_cap = bb_peq_words(big) + 8;
    std::vector<BBWinTask> tasks((size_t)cap / BB_ALIGNMENT_INTERVAL + 8), fb1(tasks.size()), fb2(tasks.size());
    int rounds = 0;
    for (; rounds < 12 && rd.status != BB_READ_DONE; rounds++) {
        int c_mut = 0, n_tasks = 0, c4 = 0, n_fb1 = 0, c8 = 0, n_fb2 = 0, cw = 0, pending = 0;
        emu::run_block(BB_MUTP_THREADS, [&]() { bb_k_mutate(B, em, seed, &c_mut, &order0, 1); });
        emu::run_warp([&]() { bb_k_window_tasks(B, &order0, 1, tasks.data(), &n_tasks); });
        emu::run_warp([&]() {
            bb_k_window_lane_hist<4, 4>(B, em, tasks.data(), &n_tasks, seed, whist.data(), ltbuf.data(), &c4, fb1.data(), &n_fb1);
        });
        emu::run_warp([&]() {
            bb_k_window_lane_hist<BB_WIN_LW, 4>(B, em, fb1.data(), &n_fb1, seed, whist.data(), ltbuf.data(), &c8, fb2.data(), &n_fb2);
        });
        emu::run_block(BB_WARPS_PER_CTA * 32, [&]() { bb_k_window_warp(B, em, pool, fb2.data(), &n_fb2, seed, &cw); });
        emu::run_warp([&]() { bb_k_replay(B, &order0, 1, k, &pending); });
    }
    if (rd.status != BB_READ_DONE) return -1;
    // K3: bb_k_join writes the read, its match bitmap and the tighter bound on its distance to the fragment
    std::vector<uint8_t> seq((size_t)rd.seq_len + 64, 0);
    std::vector<uint4> speq((size_t)bb_peq_words(rd.seq_len) + 8), speq_want(speq.size());
    B.seq = seq.data(); B.speq = speq.data();
    const int loop_upper = rd.upper;
    emu::run_block(256, [&]() { bb_k_join(B, em); });
    out8[0] = rd.loop_count; out8[1] = rd.change_count; out8[2] = rd.n_align; out8[3] = rd.seq_len;
    out8[4] = rd.start_trim; out8[5] = rd.end_trim; out8[6] = rd.upper; out8[7] = rd.flags;
    int w = 0;
    for (int x = 0; x < frag_len; x++) {
        const uint32_t st = state[(size_t)x];
        if (st == BB_SLOT_NONE) { if (w < joined_cap) joined_out[w] = fr[(size_t)x]; w++; }
        else for (int c = 0; c < (int)(st & 0xff); c++) { if (w < joined_cap) joined_out[w] = bb_slot_char(em, st, c); w++; }
    }
    if (w != rd.seq_len) return -2;
    if (std::memcmp(seq.data(), joined_out, (size_t)std::min(w, joined_cap)) != 0) return -3;   // the kernel's join
    emu::run_warp([&]() { bb_build_peq(seq.data(), rd.seq_len, speq_want.data()); });
    if (std::memcmp(speq.data(), speq_want.data(), (size_t)bb_peq_words(rd.seq_len) * sizeof(uint4)) != 0) return -4;
    if (rd.upper > loop_upper) return -5;
    if (padded_out) std::memcpy(padded_out, fr.data(), (size_t)frag_len);
    return rounds;
}


// get_qscores (qscore_model.py:32-75) under the emulator: the alignment task pipeline of emu_tasks_align, then
// bb_k_qscores_pair with the hash table bb_upload_qscore_model builds from the packed keys.  qual_out: n quality characters;
// out5 as emu_tasks_align.
extern "C" __attribute__((visibility("default")))
int emu_get_qscores(const uint8_t *seq, int n, const uint8_t *frag, int m, int upper, int kmer_size, int32_t n_keys,
                    const uint64_t *keys, const int32_t *row_off, const uint8_t *scores, const double *cum,
                    unsigned long long seed, unsigned long long read_index, uint8_t *qual_out, int *out5) {
    std::vector<uint8_t> ops((size_t)n + 8);
    std::vector<unsigned int> dcnt((size_t)n + 8);
    const int rc = emu_tasks_align(seq, n, frag, m, upper, ops.data(), dcnt.data(), out5);
    if (rc || out5[4] || out5[2]) return 1;
    uint32_t bits = 6;
    while ((1ull << bits) < 2ull * (uint64_t)n_keys) bits++;
    const size_t hsize = (size_t)1 << bits;
    std::vector<uint64_t> hk(hsize, 0);
    std::vector<int32_t> hv(hsize, -1);
    for (int32_t i = 0; i < n_keys; i++) {
        uint32_t h = (uint32_t)((keys[i] * 0x9E3779B97F4A7C15ull) >> (64 - bits));
        while (hk[h] != 0 && hk[h] != keys[i]) h = (h + 1) & (uint32_t)(hsize - 1);
        hk[h] = keys[i]; hv[h] = i;
    }
    BBQScoreModelDev qm;
    std::memset(&qm, 0, sizeof(qm));
    qm.kmer_size = kmer_size; qm.hkeys = hk.data(); qm.hvals = hv.data(); qm.hbits = bits; qm.row_off = row_off;
    qm.scores = scores; qm.cum = cum;
    emu::run_block(256, [&]() { bb_k_qscores_pair(ops.data(), dcnt.data(), n, qm, seed, read_index, qual_out); });
    return 0;
}


// The counting kernels of the model builders (bb_models.cuh) under the emulator, with the interface of
// bb_count_kmer_alternatives / bb_count_cigar_qscores (include/badread_b200.h; which = 0 / 1): one emulated CTA per alignment.
extern "C" __attribute__((visibility("default")))
int emu_count_windows(int which, int k, int max_del, int32_t n_aln, const uint8_t *read, const uint8_t *qual,
                      const int64_t *read_off, const uint8_t *ref, const int64_t *ref_off, const uint32_t *ops,
                      const int32_t *op_read0, const int32_t *op_ref0, const int64_t *ops_off, int64_t table_cap,
                      uint64_t *keys_out, uint64_t *first_out, uint32_t *counts_out, int64_t *n_entries, uint64_t *overall_out,
                      int64_t ovf_cap, int32_t *ovf_aln, int32_t *ovf_pos, int32_t *ovf_k, int64_t *n_ovf) {
    const int per_slot = which ? BBM_NQ : 1;
    BBMAln A;
    A.read = read; A.qual = qual; A.ref = ref; A.read_off = read_off; A.ref_off = ref_off; A.ops_off = ops_off; A.ops = ops;
    A.op_read0 = op_read0; A.op_ref0 = op_ref0;
    std::vector<unsigned long long> keys((size_t)table_cap, BBM_EMPTY), first((size_t)table_cap, BBM_EMPTY);
    std::vector<unsigned int> counts((size_t)table_cap * per_slot, 0u);
    int status[4] = {0, 0, 0, 0};
    unsigned long long novf = 0, n = 0;
    std::vector<unsigned long long> overall(BBM_NQ, 0ull);
    BBMTable T;
    T.keys = keys.data(); T.first = first.data(); T.counts = counts.data(); T.cap = table_cap; T.status = status;
    T.ovf_aln = ovf_aln; T.ovf_pos = ovf_pos; T.ovf_k = ovf_k; T.n_ovf = &novf; T.ovf_cap = ovf_cap;
    std::vector<int> rp((size_t)ref_off[n_aln] + 8), dc((size_t)read_off[n_aln] + 8), lead((size_t)n_aln + 8);
    std::vector<uint8_t> ism((size_t)ref_off[n_aln] + 8), sym((size_t)read_off[n_aln] + 8);
    for (int a = 0; a < n_aln; a++) {
        blockIdx.x = (unsigned)a;
        if (which) emu::run_block(256, [&]() { bbm_k_cigar_qscores(A, n_aln, k, max_del, sym.data(), dc.data(), lead.data(), T, overall.data()); });
        else emu::run_block(256, [&]() { bbm_k_kmer_alternatives(A, n_aln, k, rp.data(), ism.data(), T); });
    }
    blockIdx.x = 0;
    *n_ovf = (int64_t)novf;
    if (status[0] || status[1]) { *n_entries = 0; return -4; }
    for (long long b = 0; b < (table_cap + 255) / 256; b++) {
        blockIdx.x = (unsigned)b;
        emu::run_block(256, [&]() {
            bbm_k_compact(T, per_slot, (unsigned long long *)keys_out, (unsigned long long *)first_out, counts_out, &n, table_cap);
        });
    }
    blockIdx.x = 0;
    *n_entries = (int64_t)n;
    if (which) for (int q = 0; q < BBM_NQ; q++) overall_out[q] = overall[(size_t)q];
    return 0;
}


// K1 bb_k_build_fragments and K6 bb_k_compact for ONE read under the emulator.  The read is described like in
// bb_batch_upload: n_segs segments (kind, src, len) over the reference `ref` and the literal pool `lit`.  Outputs: the
// padded fragment, the k-mer row index per position, and a status that is 0 iff state / ctime were reset, the fragment's
// match bitmap equals bb_build_peq of the fragment, and bb_k_compact copies seq[start_trim : start_trim + out_len).
extern "C" __attribute__((visibility("default")))
int emu_build_fragment(const uint8_t *ref, const uint8_t *lit, const int32_t *seg_kind, const int64_t *seg_src,
                       const int32_t *seg_len, int n_segs, int k, const int32_t *kmer_to_row, unsigned long long seed,
                       unsigned long long read_index, uint8_t *frag_out, int32_t *kidx_out) {
    uint8_t comp[256];
    std::memset(comp, 'N', sizeof(comp));
    const char *from = "ATGCatgcRYSWKMBVDHNryswkmbvdhn.-?";
    const char *to = "TACGtacgYRSWMKVBHDNyrswmkvbhdn.-?";
    for (int i = 0; from[i]; i++) comp[(uint8_t)from[i]] = (uint8_t)to[i];
    std::memcpy(bb_c_comp, comp, 256);        // (the library: cudaMemcpyToSymbol in create_worker)
    std::vector<bb_segment> segs((size_t)n_segs);
    int len = 0;
    for (int s = 0; s < n_segs; s++) { segs[(size_t)s] = bb_segment{seg_src[s], seg_len[s], seg_kind[s]}; len += seg_len[s]; }
    const int frag_len = len + 2 * k;
    int seg_off[2] = {0, n_segs};
    BBReadDev rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.frag_len = frag_len;
    std::vector<uint8_t> fr((size_t)frag_len + 64, 0);
    std::vector<uint32_t> state((size_t)frag_len + 64, 7u);
    std::vector<unsigned int> ctime((size_t)frag_len + 64, 7u);
    std::vector<int> kidx((size_t)frag_len + 64, -7);
    std::vector<uint4> fpeq((size_t)bb_peq_words(frag_len) + 8), want(fpeq.size());
    BBBatchDev B;
    std::memset(&B, 0, sizeof(B));
    B.n_reads = 1; B.read_index = &read_index; B.seg_off = seg_off; B.segs = segs.data(); B.lit = lit; B.reads = &rd;
    B.frag = fr.data(); B.state = state.data(); B.ctime = ctime.data(); B.kidx = kidx.data(); B.fpeq = fpeq.data();
    blockIdx.x = 0;
    emu::run_block(256, [&]() { bb_k_build_fragments(B, ref, k, seed, kmer_to_row); });
    std::memcpy(frag_out, fr.data(), (size_t)frag_len);
    for (int x = 0; x + k <= frag_len; x++) kidx_out[x] = kidx[(size_t)x];
    int status = 0;
    for (int x = 0; x < frag_len; x++) if (state[(size_t)x] != BB_SLOT_NONE || ctime[(size_t)x] != 0u) status |= 1;
    emu::run_warp([&]() { bb_build_peq(fr.data(), frag_len, want.data()); });
    if (std::memcmp(fpeq.data(), want.data(), (size_t)bb_peq_words(frag_len) * sizeof(uint4)) != 0) status |= 2;
    // K6 on the same buffers: out = seq[start_trim : start_trim + out_len), qualities likewise
    std::vector<uint8_t> qual((size_t)frag_len + 64), out_seq((size_t)frag_len + 64, 0), out_qual((size_t)frag_len + 64, 0);
    for (int x = 0; x < frag_len; x++) qual[(size_t)x] = (uint8_t)(33 + x % 60);
    rd.seq_len = frag_len; rd.start_trim = std::min(k, frag_len); rd.end_trim = std::min(k, frag_len - rd.start_trim);
    rd.out_len = frag_len - rd.start_trim - rd.end_trim;
    B.seq = fr.data(); B.qual = qual.data(); B.out_seq = out_seq.data(); B.out_qual = out_qual.data();
    emu::run_block(256, [&]() { bb_k_compact(B); });
    if (std::memcmp(out_seq.data(), fr.data() + rd.start_trim, (size_t)rd.out_len) != 0 ||
        std::memcmp(out_qual.data(), qual.data() + rd.start_trim, (size_t)rd.out_len) != 0) status |= 4;
    return status;
}
