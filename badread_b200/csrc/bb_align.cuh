// bb_align.cuh — warp-level global alignment with edlib's path semantics.
//
// Replaces edlib.align(query, target, mode='NW', task='path') at the reference's call sites
// (simulate.py:330,340 — query = original window, target = mutated window; qscore_model.py:37 —
// query = mutated read, target = original fragment).  edlib (third-party, not vendored) defines the path as:
//   * traceback from the bottom-right cell preferring UP ('I', consumes a query char) over LEFT ('D') over the
//     diagonal, when its traceback state estimate 20*ceil(|q|/64)*|t| + 8*|t| is below 1 MiB;
//   * otherwise Hirschberg on the target (split at |t|/2, smallest interior query row whose left+right scores
//     equal the best score, then row -1, then row |q|-1) recursing with the same switch.
//
// One warp aligns one pair.  The DP is the Myers/Hyyrö bit-vector recurrence restricted to the Ukkonen band
// j - a <= i <= j + b.  Rows are cut into chunks of L 32-row words; chunk u belongs to lane slot (u mod K) of a
// K-lane group and works on column (step - u): a wavefront that follows the diagonal, so a pass takes about
// `ncols` steps whatever the band width (L grows with the band).  A lane retires a chunk when the band has
// moved past it and picks up chunk u + K.  Horizontal carries (and the running score, for chunk hand-over)
// travel to the next slot by one shuffle per step.  With K = 16 two independent problems (the forward and the
// reverse pass of a Hirschberg node) share a warp.  Match masks come from a per-read bitmap built once with
// ballots (bb_build_peq).  Vertical (+1) and horizontal (+1) delta words of in-band blocks are kept per column
// for leaves ("history"), so the traceback needs two bits per cell; it consumes whole diagonal runs per step
// with warp ballots.
#pragma once
#include <cstdint>

#define BB_FULL 0xffffffffu
#define BB_INF 0x3fffffff
#define BB_OP_EQ 0
#define BB_OP_X 1
#define BB_OP_I 2
#define BB_MAX_SCORE 0x3fffff  // scores travel in 22 bits of the shuffle word
#define BB_PEQ_PAD 34          // zero words in front of and behind a read's match bitmap (covers 32-word chunks)
#define BB_PEQ_BIT0 (32 * BB_PEQ_PAD)  // bit index of the read's first base

struct BBScratch {
    uint2 *hist;     // (Pv, PhRaw) per (column, block - first_block(column))
    int hist_cap;    // entries
    int8_t *hbuf;    // strip fallback: per-column horizontal delta leaving the bottom of the previous strip
    int hbuf_cap;
    int *L, *R;      // Hirschberg column scores (forward / reverse), indexed by row - row_lo
    int lr_cap;
    int *stack;      // DFS stack, 5 ints per node
    int stack_cap;   // nodes
    uint4 *peq;      // per-read match bitmap: word BB_PEQ_PAD + w holds rows [32w, 32w+32) for (A, C, G, T)
    int peq_cap;     // words
};

struct BBEmit {      // where the final alignment is written (nullptr members => counts only)
    uint8_t *ops;    // per query base: BB_OP_EQ / BB_OP_X / BB_OP_I
    unsigned int *dcnt;  // per query base: 'D' columns between this base and the next (updated atomically: leaves of
                         // one read may run concurrently and share the base at their boundary)
    int *lead_del;   // 'D' columns before the first query base
};

struct BBAlnCounts {
    int matches;     // '=' columns
    int dels;        // 'D' columns (alignment columns = query length + dels)
    int dist;        // edit distance
    int err;         // non-zero: invariant violated
};

// One banded problem handed to bb_band_pass (all lanes of a group hold the same values).
struct BBProb {
    const uint8_t *q; int qs; int n;      // query rows: row r is q[r*qs]
    const uint8_t *t; int ts; int ncols;  // target columns: column c is t[c*ts]
    int a, b;                             // band: j - a <= i <= j + b
    const uint4 *peq; int peq_bit0;       // bitmap + bit index of row 0 (BB_PEQ_BIT0 + index of that base in the
                                          // read; rows ascend for qs > 0, descend for qs < 0)
    uint2 *hist; int nb_alloc;
    int *cols_out; int cols_lo;
    uint32_t *esm = nullptr;              // SM variant only: the warp's shared-memory match-word cache, bb_esm_words(L)
    const uint4 *tpeq = nullptr; int tpeq_bit0 = 0;  // bb_band_pass_bp only: match bitmap of the read that holds the target
                                          // + bit index of column 0 in it (columns ascend for ts > 0, descend for ts < 0)
};

// Shared-memory match cache of bb_band_pass<L, ., ., true>: lane l keeps the 4 x L match words of its current chunk
// (already shifted / bit-reversed to the chunk's row order) at esm + l * (4 L + 4), component-major, so that a step
// reads its L words with L/4 conflict-free 128-bit loads (the +4 pad staggers the lanes of a quarter warp over all
// 32 banks) instead of L + 1 scattered global loads and L funnel shifts.
__host__ __device__ constexpr int bb_esm_lane_stride(int L) { return 4 * L + 4; }
__host__ __device__ constexpr int bb_esm_words(int L) { return 32 * bb_esm_lane_stride(L); }

__device__ __forceinline__ void bb_band(int n, int m, int k, int &a, int &b) {
    // a path of cost <= k from (0,0) to (n,m) has at most (k-(n-m))/2 'D' and (k+(n-m))/2 'I' moves:
    // every cell (i,j) it can visit satisfies j - a <= i <= j + b
    a = (k - (n - m)) / 2; if (a < 0) a = 0;
    b = (k + (n - m)) / 2; if (b < 0) b = 0;
    // the wavefront hands a chunk's starting score over from the chunk above, which must still be inside the band
    // at that column: that needs a + b >= 1 (a wider band is always valid)
    if (a + b < 1) b = 1;
}

__device__ __forceinline__ bool bb_uses_traceback(int n, int m) {
    return 20ll * ((n + 63) / 64) * m + 8ll * m < 1048576ll;
}

__device__ __forceinline__ int bb_first_block(int j, int a, int nblk) {
    int lo = j - a; if (lo < 0) lo = 0;
    int f = lo >> 5; if (f > nblk - 1) f = nblk - 1;
    return f;
}
__device__ __forceinline__ int bb_last_block(int j, int b, int n) {
    int hi = j + b; if (hi > n - 1) hi = n - 1;
    return hi >> 5;
}

// Match bitmap of a whole read: peq[BB_PEQ_PAD + w] = ballots of (q[32w + lane] == A/C/G/T); BB_PEQ_PAD zero words
// in front and behind, so that any window of up to 32 words that touches the read can be cut out with funnel
// shifts without bounds checks.  Size: bb_peq_words(n).
__host__ __device__ __forceinline__ int bb_peq_words(int n) { return ((n + 31) >> 5) + 2 * BB_PEQ_PAD; }

static __device__ void bb_build_peq(const uint8_t *q, int n, uint4 *peq) {
    const int lane = threadIdx.x & 31;
    const int nw = (n + 31) >> 5;
    for (int w = lane; w < BB_PEQ_PAD; w += 32) {
        peq[w] = make_uint4(0u, 0u, 0u, 0u);
        peq[BB_PEQ_PAD + nw + w] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int w = 0; w < nw; w++) {
        const int row = 32 * w + lane;
        const uint8_t c = row < n ? q[row] : 0;
        const uint32_t mA = __ballot_sync(BB_FULL, c == 'A'), mC = __ballot_sync(BB_FULL, c == 'C');
        const uint32_t mG = __ballot_sync(BB_FULL, c == 'G'), mT = __ballot_sync(BB_FULL, c == 'T');
        if (lane == 0) peq[BB_PEQ_PAD + w] = make_uint4(mA, mC, mG, mT);
    }
    __syncwarp();
}

// The four match masks of the 32 rows [R, R+32) of a problem (rows >= n are cleared).
__device__ __forceinline__ void bb_fetch_peq(const BBProb &P, int R, uint32_t &mA, uint32_t &mC, uint32_t &mG,
                                             uint32_t &mT) {
    mA = mC = mG = mT = 0u;
    const int valid = P.n - R;
    if (valid <= 0) return;
    const int s = P.qs > 0 ? P.peq_bit0 + R : P.peq_bit0 - R - 31;
    const int idx = s >> 5, sh = s & 31;
    const uint4 lo = P.peq[idx], hi = P.peq[idx + 1];
    mA = __funnelshift_r(lo.x, hi.x, sh); mC = __funnelshift_r(lo.y, hi.y, sh);
    mG = __funnelshift_r(lo.z, hi.z, sh); mT = __funnelshift_r(lo.w, hi.w, sh);
    if (P.qs < 0) { mA = __brev(mA); mC = __brev(mC); mG = __brev(mG); mT = __brev(mT); }
    if (valid < 32) {
        const uint32_t keep = (1u << valid) - 1u;
        mA &= keep; mC &= keep; mG &= keep; mT &= keep;
    }
}

// S = A + B over L 32-bit words, least significant word first (one carry chain).
template <int L>
__device__ __forceinline__ void bb_add_words(const uint32_t (&A)[L], const uint32_t (&B)[L], uint32_t (&S)[L]) {
#ifdef __CUDA_ARCH__
    if (L == 1) { S[0] = A[0] + B[0]; return; }
    asm("add.cc.u32 %0, %1, %2;" : "=r"(S[0]) : "r"(A[0]), "r"(B[0]));
#pragma unroll
    for (int x = 1; x < L - 1; x++) asm("addc.cc.u32 %0, %1, %2;" : "=r"(S[x]) : "r"(A[x]), "r"(B[x]));
    asm("addc.u32 %0, %1, %2;" : "=r"(S[L - 1]) : "r"(A[L - 1]), "r"(B[L - 1]));
#else
    uint64_t carry = 0;
    for (int x = 0; x < L; x++) {
        const uint64_t v = (uint64_t)A[x] + B[x] + carry;
        S[x] = (uint32_t)v;
        carry = v >> 32;
    }
#endif
}

// Banded NW of one or two problems (K = 32: one problem; K = 16: lanes 0-15 and 16-31 hold their own BBProb).
// Requires 1 <= a + b and (a + b) / (32 L) + 2 <= K.  HIST: store (Pv, PhRaw) of in-band blocks.  COLS: write
// D[row][ncols-1] of the in-band rows of the last column to cols_out[row - cols_lo].  Returns D[n-1][ncols-1] of
// the lane's own problem (BB_INF if the band does not contain that cell).  Values are exact for every cell on a
// path of cost <= the k the band was derived from and upper bounds elsewhere.
// A chunk is handled as ONE 32L-bit Myers word: the only cross-word dependencies of a step are the carry chain
// of the addition and the one-bit shifts, everything else is independent per word.
template <int L, bool HIST, bool COLS, bool SM = false>
__device__ int bb_band_pass(const BBProb &P, int K) {
    const int lane = threadIdx.x & 31;
    const int slot = lane & (K - 1);
    const int prev = (lane & ~(K - 1)) | ((slot + K - 1) & (K - 1));
    constexpr int CH = 32 * L;
    const int n = P.n, ncols = P.ncols, a = P.a, b = P.b, ts = P.ts;
    const int nb_alloc = P.nb_alloc;
    uint2 *const hist = P.hist;
    const int nblk = (n + 31) >> 5;
    int ulast = -1;
    if (ncols > 0 && n > 0) {
        ulast = (ncols - 1 + b) / CH;
        const int nchunks = (n + CH - 1) / CH;
        if (ulast > nchunks - 1) ulast = nchunks - 1;
    }
    const int T = __reduce_max_sync(BB_FULL, ulast >= 0 ? ncols + ulast : 0);
    const int cols_hi = min(n - 1, ncols - 1 + b);
    // Vertical deltas live in registers.  Match words: for L <= 4 the four per-letter masks of the chunk are held in
    // registers as well; wider chunks stream the words of the current target base from the (L1-resident) bitmap
    // every step (L + 1 loads and L funnel shifts) so that the register footprint stays at 2L.
    constexpr bool STREAM = (L >= 8);
    constexpr int LR = STREAM ? 1 : L;
    uint32_t Pv[L], Mv[L], eA[LR], eC[LR], eG[LR], eT[LR];
#pragma unroll
    for (int x = 0; x < L; x++) { Pv[x] = ~0u; Mv[x] = 0u; }
#pragma unroll
    for (int x = 0; x < LR; x++) { eA[x] = eC[x] = eG[x] = eT[x] = 0u; }
    const uint32_t *const ewords = reinterpret_cast<const uint32_t *>(P.peq);
    const bool fwd = P.qs > 0;
    uint32_t *const esm = (SM && STREAM) ? P.esm + lane * bb_esm_lane_stride(L) : nullptr;
    int eidx = 0, esh = 0, evalid = 0;  // bitmap word / shift of the chunk's first word; rows of the chunk below n
    int u = slot;
    int cs = max(0, CH * u - b), ce = min(ncols - 1, CH * u + CH - 1 + a);
    int ce_up = min(ncols - 1, CH * u - 1 + a);  // last column of the chunk above
    int score = 0, result = BB_INF;
    uint32_t outpack = 0;
    uint32_t tcn = 0;
    const uint8_t *tp = P.t - (long long)u * ts;  // tp + tau*ts is this lane's column at step tau
    if (u <= ulast && 0 - u >= cs && 0 - u <= ce) tcn = *tp;
    for (int tau = 0; tau < T; tau++) {
        const uint32_t in = __shfl_sync(BB_FULL, outpack, prev);
        const int c = tau - u;
        const bool active = (u <= ulast) && c >= cs && c <= ce;
        if (active) {
            const uint32_t tc = tcn;
            int hin = 1;
            if (u > 0 && c <= ce_up) hin = (int)((in >> 22) & 3u) - 1;
            if (c == cs) {  // a chunk entering the band starts from the all-(+1) upper bound below chunk u-1
                const int base = (u == 0) ? cs : (int)(in & BB_MAX_SCORE) - hin;
                score = base + CH;
#pragma unroll
                for (int x = 0; x < L; x++) { Pv[x] = ~0u; Mv[x] = 0u; }
                if (STREAM) {
                    const int s0 = fwd ? P.peq_bit0 + u * CH : P.peq_bit0 - u * CH - 31;
                    eidx = s0 >> 5; esh = s0 & 31;
                    evalid = n - u * CH;
                    if (SM) {  // cut the chunk's match words out of the bitmap once; the steps read them from smem
                        uint4 w0 = P.peq[fwd ? eidx : eidx + 1];
#pragma unroll 1
                        for (int x = 0; x < L; x++) {
                            const uint4 w1 = P.peq[fwd ? eidx + x + 1 : eidx - x];
                            uint4 e;
                            if (fwd) {
                                e.x = __funnelshift_r(w0.x, w1.x, esh); e.y = __funnelshift_r(w0.y, w1.y, esh);
                                e.z = __funnelshift_r(w0.z, w1.z, esh); e.w = __funnelshift_r(w0.w, w1.w, esh);
                            } else {
                                e.x = __brev(__funnelshift_r(w1.x, w0.x, esh)); e.y = __brev(__funnelshift_r(w1.y, w0.y, esh));
                                e.z = __brev(__funnelshift_r(w1.z, w0.z, esh)); e.w = __brev(__funnelshift_r(w1.w, w0.w, esh));
                            }
                            const int v = evalid - 32 * x;
                            const uint32_t keep = v >= 32 ? ~0u : (v <= 0 ? 0u : ((1u << v) - 1u));
                            esm[x] = e.x & keep; esm[L + x] = e.y & keep; esm[2 * L + x] = e.z & keep; esm[3 * L + x] = e.w & keep;
                            w0 = w1;
                        }
                    }
                } else {
#pragma unroll
                    for (int x = 0; x < LR; x++) bb_fetch_peq(P, u * CH + 32 * x, eA[x], eC[x], eG[x], eT[x]);
                }
            }
            const uint32_t code = (tc >> 1) & 3u;  // A->0, C->1, T->2, G->3
            const bool acgt = ((0x47544341u >> (8 * code)) & 0xffu) == tc;
            uint32_t Eq[L], Xv[L], A[L], S[L];
            if (STREAM && SM) {
                const uint4 *ep = reinterpret_cast<const uint4 *>(esm + (code ^ (code >> 1)) * L);
#pragma unroll
                for (int x = 0; x < L; x += 4) {
                    const uint4 e = ep[x >> 2];
                    if (x + 3 < L) { Eq[x] = e.x; Eq[x + 1] = e.y; Eq[x + 2] = e.z; Eq[x + 3] = e.w; }
                }
            } else if (STREAM) {
                const uint32_t *ep = ewords + (code ^ (code >> 1));  // uint4 component: A, C, G, T
                uint32_t wv[L + 1];
                if (fwd) {
#pragma unroll
                    for (int x = 0; x <= L; x++) wv[x] = ep[(eidx + x) * 4];
#pragma unroll
                    for (int x = 0; x < L; x++) Eq[x] = __funnelshift_r(wv[x], wv[x + 1], esh);
                } else {
#pragma unroll
                    for (int x = 0; x <= L; x++) wv[x] = ep[(eidx + 1 - x) * 4];
#pragma unroll
                    for (int x = 0; x < L; x++) Eq[x] = __brev(__funnelshift_r(wv[x + 1], wv[x], esh));
                }
                if (evalid < CH) {  // the chunk reaches past the node's last row: those rows match nothing
#pragma unroll
                    for (int x = 0; x < L; x++) {
                        const int v = evalid - 32 * x;
                        Eq[x] = v >= 32 ? Eq[x] : (v <= 0 ? 0u : (Eq[x] & ((1u << v) - 1u)));
                    }
                }
            } else {
#pragma unroll
                for (int x = 0; x < L; x++) {
                    const int y = x < LR ? x : 0;
                    Eq[x] = (code & 2u) ? ((code & 1u) ? eG[y] : eT[y]) : ((code & 1u) ? eC[y] : eA[y]);
                }
            }
            if (!acgt) {  // non-ACGT target character: exact byte equality against every row of the chunk
#pragma unroll
                for (int x = 0; x < L; x++) {
                    Eq[x] = 0u;
                    const int row0 = u * CH + 32 * x;
                    for (int r = 0; r < 32; r++)
                        if (row0 + r < n && P.q[(long long)(row0 + r) * P.qs] == tc) Eq[x] |= 1u << r;
                }
            }
            const uint32_t hin_neg = hin < 0 ? 1u : 0u;
#pragma unroll
            for (int x = 0; x < L; x++) Xv[x] = Eq[x] | Mv[x];
            Eq[0] |= hin_neg;
#pragma unroll
            for (int x = 0; x < L; x++) A[x] = Eq[x] & Pv[x];
            bb_add_words<L>(A, Pv, S);
            uint32_t Ph[L], Mh[L];
#pragma unroll
            for (int x = 0; x < L; x++) {
                const uint32_t Xh = (S[x] ^ Pv[x]) | Eq[x];
                Ph[x] = Mv[x] | ~(Xh | Pv[x]);
                Mh[x] = Pv[x] & Xh;
            }
            const int hout = (int)(Ph[L - 1] >> 31) - (int)(Mh[L - 1] >> 31);
            if (HIST) {
#pragma unroll
                for (int x = 0; x < L; x++) S[x] = Ph[x];  // PhRaw of this column (S is free now)
            }
#pragma unroll
            for (int x = L - 1; x >= 0; x--) {
                const uint32_t ph_lo = x > 0 ? Ph[x - 1] : (hin > 0 ? 0x80000000u : 0u);
                const uint32_t mh_lo = x > 0 ? Mh[x - 1] : (hin_neg << 31);
                const uint32_t phs = __funnelshift_l(ph_lo, Ph[x], 1);
                const uint32_t mhs = __funnelshift_l(mh_lo, Mh[x], 1);
                Pv[x] = mhs | ~(Xv[x] | phs);
                Mv[x] = phs & Xv[x];
            }
            if (HIST) {
                const int bf = bb_first_block(c, a, nblk), bl = bb_last_block(c, b, n);
#pragma unroll
                for (int x = 0; x < L; x++) {
                    const int blk = u * L + x;
                    const int rel = blk - bf;
                    if (rel >= 0 && rel < nb_alloc && blk <= bl) hist[c * nb_alloc + rel] = make_uint2(Pv[x], S[x]);
                }
            }
            score += hout;
            outpack = ((uint32_t)(hout + 1) << 22) | ((uint32_t)score & BB_MAX_SCORE);
            if (c == ncols - 1) {
                int run = score;
#pragma unroll
                for (int x = L - 1; x >= 0; x--) {
                    const int row0 = u * CH + 32 * x;
                    if (COLS) {
                        int rr = run;
                        for (int r = 31; r >= 0; r--) {
                            const int row = row0 + r;
                            if (row < n && row >= P.cols_lo && row <= cols_hi) P.cols_out[row - P.cols_lo] = rr;
                            rr -= (int)((Pv[x] >> r) & 1u) - (int)((Mv[x] >> r) & 1u);
                        }
                    }
                    if (row0 <= n - 1 && n - 1 < row0 + 32) {
                        const int bit = (n - 1) - row0;
                        const uint32_t up = bit == 31 ? 0u : (Pv[x] >> (bit + 1));
                        const uint32_t um = bit == 31 ? 0u : (Mv[x] >> (bit + 1));
                        result = run - __popc(up) + __popc(um);
                    }
                    run -= __popc(Pv[x]) - __popc(Mv[x]);
                }
            }
            if (c == ce) {  // the band has moved past this chunk: take over chunk u + K
                u += K;
                cs = max(0, CH * u - b);
                ce = min(ncols - 1, CH * u + CH - 1 + a);
                ce_up = min(ncols - 1, CH * u - 1 + a);
                tp -= (long long)K * ts;
            }
        }
        const int cn = tau + 1 - u;
        if (u <= ulast && cn >= cs && cn <= ce) tcn = tp[(long long)(tau + 1) * ts];
    }
    const int owner = (lane & ~(K - 1)) | ((n > 0 ? (n - 1) / CH : 0) & (K - 1));
    result = __shfl_sync(BB_FULL, result, owner);
    __syncwarp();
    return result;
}

// The wavefront of bb_band_pass spread over the WARPS warps of a CTA group (K = 32 WARPS lane slots, one problem):
// a band that needs L words per lane on one warp needs L / WARPS here, so a step is WARPS times shorter and the
// warps issue on different schedulers.  Distance only, column scores out (the node passes of the Hirschberg recursion).
// Lane 31 of a warp hands its chunk's horizontal deltas to lane 0 of the next warp through a ring of mailboxes in
// shared memory: mbox[wg * 8 + (step & 7)] holds warp wg's output of that step and progress[wg] the number of steps it
// has published.  Lane 0 polls its predecessor's progress (bounded: a warp that waits longer than any correct run can
// take gives up with an error instead of hanging); no CTA barrier is involved, the warps drift apart by at most WARPS - 1
// steps because the dependency runs in a ring.  All warps of the group must call this with the same problem; `wg` is
// the warp's index in the group; progress[] must be zero on entry.  Returns 0, or 256 if the hand-over timed out.
template <int L, int WARPS>
__device__ int bb_band_pass_mw(const BBProb &P, int wg, volatile uint32_t *mbox, volatile int *progress) {
    constexpr int K = 32 * WARPS;
    constexpr int CH = 32 * L;
    const int lane = threadIdx.x & 31;
    const int slot = wg * 32 + lane;
    const int n = P.n, ncols = P.ncols, a = P.a, b = P.b, ts = P.ts;
    int ulast = -1;
    if (ncols > 0 && n > 0) {
        ulast = (ncols - 1 + b) / CH;
        const int nchunks = (n + CH - 1) / CH;
        if (ulast > nchunks - 1) ulast = nchunks - 1;
    }
    const int T = ulast >= 0 ? ncols + ulast : 0;  // the same for every lane of the group
    const int cols_hi = min(n - 1, ncols - 1 + b);
    constexpr bool STREAM = (L >= 8);
    constexpr int LR = STREAM ? 1 : L;
    uint32_t Pv[L], Mv[L], eA[LR], eC[LR], eG[LR], eT[LR];
#pragma unroll
    for (int x = 0; x < L; x++) { Pv[x] = ~0u; Mv[x] = 0u; }
#pragma unroll
    for (int x = 0; x < LR; x++) { eA[x] = eC[x] = eG[x] = eT[x] = 0u; }
    const bool fwd = P.qs > 0;
    uint32_t *const esm = STREAM ? P.esm + lane * bb_esm_lane_stride(L) : nullptr;
    int u = slot;
    int cs = max(0, CH * u - b), ce = min(ncols - 1, CH * u + CH - 1 + a);
    int ce_up = min(ncols - 1, CH * u - 1 + a);  // last column of the chunk above
    int score = 0;
    uint32_t outpack = 0;
    uint32_t tcn = 0;
    const uint8_t *tp = P.t - (long long)u * ts;  // tp + tau*ts is this lane's column at step tau
    if (u <= ulast && 0 - u >= cs && 0 - u <= ce) tcn = *tp;
    const int src = (wg + WARPS - 1) % WARPS;
    int failed = 0;
    for (int tau = 0; tau < T; tau++) {
        // the previous slot's output of step tau - 1: from the lane below, or from the warp before through its mailbox
        uint32_t in = __shfl_up_sync(BB_FULL, outpack, 1);
        if (lane == 0 && tau > 0) {
            int spins = 0;
            while (!failed && progress[src] < tau) {  // (after one timeout the pass no longer waits: it ends, flagged)
#ifdef BB_EMULATOR
                emu::yield();
#endif
                if (++spins > (1 << 24)) { failed = 256; break; }
            }
            in = mbox[src * 8 + ((tau - 1) & 7)];
        }
        const int c = tau - u;
        const bool active = (u <= ulast) && c >= cs && c <= ce;
        if (active) {
            const uint32_t tc = tcn;
            int hin = 1;
            if (u > 0 && c <= ce_up) hin = (int)((in >> 22) & 3u) - 1;
            if (c == cs) {  // a chunk entering the band starts from the all-(+1) upper bound below chunk u-1
                const int base = (u == 0) ? cs : (int)(in & BB_MAX_SCORE) - hin;
                score = base + CH;
#pragma unroll
                for (int x = 0; x < L; x++) { Pv[x] = ~0u; Mv[x] = 0u; }
                if (STREAM) {  // cut the chunk's match words out of the bitmap once; the steps read them from smem
                    const int s0 = fwd ? P.peq_bit0 + u * CH : P.peq_bit0 - u * CH - 31;
                    const int eidx = s0 >> 5, esh = s0 & 31;
                    const int evalid = n - u * CH;
                    uint4 w0 = P.peq[fwd ? eidx : eidx + 1];
#pragma unroll 1
                    for (int x = 0; x < L; x++) {
                        const uint4 w1 = P.peq[fwd ? eidx + x + 1 : eidx - x];
                        uint4 e;
                        if (fwd) {
                            e.x = __funnelshift_r(w0.x, w1.x, esh); e.y = __funnelshift_r(w0.y, w1.y, esh);
                            e.z = __funnelshift_r(w0.z, w1.z, esh); e.w = __funnelshift_r(w0.w, w1.w, esh);
                        } else {
                            e.x = __brev(__funnelshift_r(w1.x, w0.x, esh)); e.y = __brev(__funnelshift_r(w1.y, w0.y, esh));
                            e.z = __brev(__funnelshift_r(w1.z, w0.z, esh)); e.w = __brev(__funnelshift_r(w1.w, w0.w, esh));
                        }
                        const int v = evalid - 32 * x;
                        const uint32_t keep = v >= 32 ? ~0u : (v <= 0 ? 0u : ((1u << v) - 1u));
                        esm[x] = e.x & keep; esm[L + x] = e.y & keep; esm[2 * L + x] = e.z & keep; esm[3 * L + x] = e.w & keep;
                        w0 = w1;
                    }
                } else {
#pragma unroll
                    for (int x = 0; x < LR; x++) bb_fetch_peq(P, u * CH + 32 * x, eA[x], eC[x], eG[x], eT[x]);
                }
            }
            const uint32_t code = (tc >> 1) & 3u;  // A->0, C->1, T->2, G->3
            const bool acgt = ((0x47544341u >> (8 * code)) & 0xffu) == tc;
            uint32_t Eq[L], Xv[L], A[L], S[L];
            if (STREAM) {
                const uint4 *ep = reinterpret_cast<const uint4 *>(esm + (code ^ (code >> 1)) * L);
#pragma unroll
                for (int x = 0; x < L; x += 4) {
                    const uint4 e = ep[x >> 2];
                    if (x + 3 < L) { Eq[x] = e.x; Eq[x + 1] = e.y; Eq[x + 2] = e.z; Eq[x + 3] = e.w; }
                }
            } else {
#pragma unroll
                for (int x = 0; x < L; x++) {
                    const int y = x < LR ? x : 0;
                    Eq[x] = (code & 2u) ? ((code & 1u) ? eG[y] : eT[y]) : ((code & 1u) ? eC[y] : eA[y]);
                }
            }
            if (!acgt) {  // non-ACGT target character: exact byte equality against every row of the chunk
#pragma unroll
                for (int x = 0; x < L; x++) {
                    Eq[x] = 0u;
                    const int row0 = u * CH + 32 * x;
                    for (int r = 0; r < 32; r++)
                        if (row0 + r < n && P.q[(long long)(row0 + r) * P.qs] == tc) Eq[x] |= 1u << r;
                }
            }
            const uint32_t hin_neg = hin < 0 ? 1u : 0u;
#pragma unroll
            for (int x = 0; x < L; x++) Xv[x] = Eq[x] | Mv[x];
            Eq[0] |= hin_neg;
#pragma unroll
            for (int x = 0; x < L; x++) A[x] = Eq[x] & Pv[x];
            bb_add_words<L>(A, Pv, S);
            uint32_t Ph[L], Mh[L];
#pragma unroll
            for (int x = 0; x < L; x++) {
                const uint32_t Xh = (S[x] ^ Pv[x]) | Eq[x];
                Ph[x] = Mv[x] | ~(Xh | Pv[x]);
                Mh[x] = Pv[x] & Xh;
            }
            const int hout = (int)(Ph[L - 1] >> 31) - (int)(Mh[L - 1] >> 31);
#pragma unroll
            for (int x = L - 1; x >= 0; x--) {
                const uint32_t ph_lo = x > 0 ? Ph[x - 1] : (hin > 0 ? 0x80000000u : 0u);
                const uint32_t mh_lo = x > 0 ? Mh[x - 1] : (hin_neg << 31);
                const uint32_t phs = __funnelshift_l(ph_lo, Ph[x], 1);
                const uint32_t mhs = __funnelshift_l(mh_lo, Mh[x], 1);
                Pv[x] = mhs | ~(Xv[x] | phs);
                Mv[x] = phs & Xv[x];
            }
            score += hout;
            outpack = ((uint32_t)(hout + 1) << 22) | ((uint32_t)score & BB_MAX_SCORE);
            if (c == ncols - 1) {
                int run = score;
#pragma unroll
                for (int x = L - 1; x >= 0; x--) {
                    const int row0 = u * CH + 32 * x;
                    int rr = run;
                    for (int r = 31; r >= 0; r--) {
                        const int row = row0 + r;
                        if (row < n && row >= P.cols_lo && row <= cols_hi) P.cols_out[row - P.cols_lo] = rr;
                        rr -= (int)((Pv[x] >> r) & 1u) - (int)((Mv[x] >> r) & 1u);
                    }
                    run -= __popc(Pv[x]) - __popc(Mv[x]);
                }
            }
            if (c == ce) {  // the band has moved past this chunk: take over chunk u + K
                u += K;
                cs = max(0, CH * u - b);
                ce = min(ncols - 1, CH * u + CH - 1 + a);
                ce_up = min(ncols - 1, CH * u - 1 + a);
                tp -= (long long)K * ts;
            }
        }
        const int cn = tau + 1 - u;
        if (u <= ulast && cn >= cs && cn <= ce) tcn = tp[(long long)(tau + 1) * ts];
        if (lane == 31) {  // publish this step's output for lane 0 of the next warp
            mbox[wg * 8 + (tau & 7)] = outpack;
            __threadfence_block();
            progress[wg] = tau + 1;
        }
    }
    __syncwarp();
    return __reduce_or_sync(BB_FULL, (unsigned)failed);
}

// Bit planes of the 32 target columns [c, c + 32) of a problem, bit j = column c + j: the 2-bit code of the base
// (A 00, C 01, T 10, G 11 - the code (char >> 1) & 3 of the other passes) and whether it is one of ACGT at all.  Cut out of
// the match bitmap of the read that holds the target, like bb_fetch_peq cuts the query rows out of theirs.
__device__ __forceinline__ void bb_fetch_target_planes(const BBProb &P, int c, uint32_t &b0, uint32_t &b1, uint32_t &ok) {
    const int s = P.ts > 0 ? P.tpeq_bit0 + c : P.tpeq_bit0 - c - 31;
    const int idx = s >> 5, sh = s & 31;
    const uint4 lo = P.tpeq[idx], hi = P.tpeq[idx + 1];
    uint32_t mA = __funnelshift_r(lo.x, hi.x, sh), mC = __funnelshift_r(lo.y, hi.y, sh);
    uint32_t mG = __funnelshift_r(lo.z, hi.z, sh), mT = __funnelshift_r(lo.w, hi.w, sh);
    if (P.ts < 0) { mA = __brev(mA); mC = __brev(mC); mG = __brev(mG); mT = __brev(mT); }
    b0 = mC | mG; b1 = mT | mG; ok = mA | mC | mG | mT;
}

// The node passes of the lean warp kernels: distance only, column scores out (the outputs of bb_band_pass<L, false, true>),
// TWO columns per wavefront step - chunk u works on columns 2 (step - u) and 2 (step - u) + 1; one shuffle carries the two
// horizontal deltas of the chunk above (2 bits each) and its score after the first of those columns - on BIT PLANES.  A
// step whose columns are interior to the chunk (not its first, not its last, all ACGT) runs two bare Myers steps;
// everything else takes the general per-column path.  What a wavefront step spends around the Myers recurrence itself is
// most of it for narrow chunks (~32 instructions per column against 11 L + 6), so:
//   * the target is not fetched byte by byte: the codes of 32 columns are two words cut out of the target's match bitmap
//     (one refill per 16 steps instead of two byte loads with their address arithmetic and ACGT tests per step);
//   * the chunk keeps its rows as two code planes q0, q1 instead of four letter masks: the match word of a column is
//     ~((q0 ^ c0) | (q1 ^ c1)) with the column's code bits spread to words - two logic operations per word instead of a
//     three-way select, 2 L fewer registers;
//   * rows past the node's last row need no masking (they are rows of 'A' appended to the query: the recurrence only
//     carries information downwards, and the corner / column scores subtract the vertical deltas below the last row).
// Exactness for anything that is not ACGT: a target column outside ACGT (its `ok` bit is clear) and a chunk whose rows are
// not all ACGT (qok) take the per-column path with the letter masks fetched again and exact byte comparison.
// Requires P.tpeq / P.tpeq_bit0.
template <int L>
__device__ int bb_band_pass_bp(const BBProb &P, int K) {
    constexpr int CB = 2;
    const int lane = threadIdx.x & 31;
    const int slot = lane & (K - 1);
    const int prev = (lane & ~(K - 1)) | ((slot + K - 1) & (K - 1));
    constexpr int CH = 32 * L;
    const int n = P.n, ncols = P.ncols, a = P.a, b = P.b;
    int ulast = -1;
    if (ncols > 0 && n > 0) {
        ulast = (ncols - 1 + b) / CH;
        const int nchunks = (n + CH - 1) / CH;
        if (ulast > nchunks - 1) ulast = nchunks - 1;
    }
    const int T = __reduce_max_sync(BB_FULL, ulast >= 0 ? (ncols - 1) / CB + ulast + 1 : 0);
    const int cols_hi = min(n - 1, ncols - 1 + b);
    const bool even_band = ((a | b) & 1) == 0;
    uint32_t Pv[L], Mv[L], q0[L], q1[L], q0n[L], q1n[L];
#pragma unroll
    for (int x = 0; x < L; x++) { Pv[x] = ~0u; Mv[x] = 0u; q0[x] = q1[x] = q0n[x] = q1n[x] = 0u; }
    bool qok = true;   // every row of the current chunk that exists is one of ACGT
    bool qokn = true;  // ... of the lane's NEXT chunk, whose planes (q0n, q1n) are fetched a chunk ahead: when a lane
                       // takes a chunk the other 31 wait for it, so that moment should not begin with loads
    auto fetch_chunk = [&](int uu) {
        qokn = true;
#pragma unroll
        for (int x = 0; x < L; x++) {
            uint32_t eA, eC, eG, eT;
            bb_fetch_peq(P, uu * CH + 32 * x, eA, eC, eG, eT);
            q0n[x] = eC | eG; q1n[x] = eT | eG;
            const int v = n - (uu * CH + 32 * x);
            const uint32_t rows = v >= 32 ? ~0u : (v <= 0 ? 0u : ((1u << v) - 1u));
            qokn = qokn && (((eA | eC | eG | eT) | ~rows) == ~0u);
        }
    };
    int u = slot;
    if (u <= ulast) fetch_chunk(u);
    int cs = max(0, CH * u - b), ce = min(ncols - 1, CH * u + CH - 1 + a);
    int ce_up = u == 0 ? -1 : min(ncols - 1, CH * u - 1 + a);  // last column of the chunk above (none above chunk 0)
    int score = 0, result = BB_INF;
    uint32_t outpack = 0;
    uint32_t tb0 = 0, tb1 = 0, tok = 0;  // code planes / ACGT flags of the target columns [tbase, tbase + 32)
    int tbase = -(1 << 28);
    // one Myers step of the whole chunk on match words Eq with horizontal input hin; returns the horizontal output
    auto myers = [&](uint32_t (&Eq)[L], int hin) -> int {
        uint32_t Xv[L], A[L], S[L], Ph[L], Mh[L];
        const uint32_t hin_neg = hin < 0 ? 1u : 0u;
#pragma unroll
        for (int x = 0; x < L; x++) Xv[x] = Eq[x] | Mv[x];
        Eq[0] |= hin_neg;
#pragma unroll
        for (int x = 0; x < L; x++) A[x] = Eq[x] & Pv[x];
        bb_add_words<L>(A, Pv, S);
#pragma unroll
        for (int x = 0; x < L; x++) {
            const uint32_t Xh = (S[x] ^ Pv[x]) | Eq[x];
            Ph[x] = Mv[x] | ~(Xh | Pv[x]);
            Mh[x] = Pv[x] & Xh;
        }
        const int hout = (int)(Ph[L - 1] >> 31) - (int)(Mh[L - 1] >> 31);
#pragma unroll
        for (int x = L - 1; x >= 0; x--) {
            const uint32_t ph_lo = x > 0 ? Ph[x - 1] : (hin > 0 ? 0x80000000u : 0u);
            const uint32_t mh_lo = x > 0 ? Mh[x - 1] : (hin_neg << 31);
            const uint32_t phs = __funnelshift_l(ph_lo, Ph[x], 1);
            const uint32_t mhs = __funnelshift_l(mh_lo, Mh[x], 1);
            Pv[x] = mhs | ~(Xv[x] | phs);
            Mv[x] = phs & Xv[x];
        }
        return hout;
    };
    // match words of a column whose code bits are (k0, k1), from the chunk's planes
    auto planes = [&](uint32_t k0, uint32_t k1, uint32_t (&Eq)[L]) {
        const uint32_t m0 = 0u - (k0 & 1u), m1 = 0u - (k1 & 1u);
#pragma unroll
        for (int x = 0; x < L; x++) Eq[x] = ~((q0[x] ^ m0) | (q1[x] ^ m1));
    };
    // any column of the chunk (its first, its last, non-ACGT characters, the last column of the pass)
    // a chunk entering the band (at its column cs) starts from the all-(+1) upper bound below chunk u-1
    auto enter_chunk = [&](int hin, int above) {
        score = ((u == 0) ? cs : above - hin) + CH;
        qok = qokn;
#pragma unroll
        for (int x = 0; x < L; x++) { Pv[x] = ~0u; Mv[x] = 0u; q0[x] = q0n[x]; q1[x] = q1n[x]; }
        if (u + K <= ulast) fetch_chunk(u + K);  // not needed before the band has passed this chunk
    };
    auto column = [&](int c, uint32_t k0, uint32_t k1, bool tplain, int hin, int above, bool entered) -> int {
        if (c == cs && !entered) enter_chunk(hin, above);
        uint32_t Eq[L];
        if (qok && tplain) planes(k0, k1, Eq);
        else {  // exact: the letter masks again; a target character outside ACGT is compared byte by byte
            const uint32_t code = (k0 & 1u) | ((k1 & 1u) << 1);
#pragma unroll
            for (int x = 0; x < L; x++) {
                uint32_t eA, eC, eG, eT;
                bb_fetch_peq(P, u * CH + 32 * x, eA, eC, eG, eT);
                Eq[x] = (code & 2u) ? ((code & 1u) ? eG : eT) : ((code & 1u) ? eC : eA);
                if (!tplain) {
                    const uint32_t tc = P.t[(long long)c * P.ts];
                    Eq[x] = 0u;
                    const int row0 = u * CH + 32 * x;
                    for (int r = 0; r < 32; r++)
                        if (row0 + r < n && P.q[(long long)(row0 + r) * P.qs] == tc) Eq[x] |= 1u << r;
                }
            }
        }
        const int hout = myers(Eq, hin);
        score += hout;
        if (c == ncols - 1) {
            int run = score;
#pragma unroll
            for (int x = L - 1; x >= 0; x--) {
                const int row0 = u * CH + 32 * x;
                int rr = run;
                for (int r = 31; r >= 0; r--) {
                    const int row = row0 + r;
                    if (row < n && row >= P.cols_lo && row <= cols_hi) P.cols_out[row - P.cols_lo] = rr;
                    rr -= (int)((Pv[x] >> r) & 1u) - (int)((Mv[x] >> r) & 1u);
                }
                if (row0 <= n - 1 && n - 1 < row0 + 32) {
                    const int bit = (n - 1) - row0;
                    const uint32_t up = bit == 31 ? 0u : (Pv[x] >> (bit + 1));
                    const uint32_t um = bit == 31 ? 0u : (Mv[x] >> (bit + 1));
                    result = run - __popc(up) + __popc(um);
                }
                run -= __popc(Pv[x]) - __popc(Mv[x]);
            }
        }
        return hout;
    };
    for (int s = 0; s < T; s++) {
        const uint32_t in = __shfl_sync(BB_FULL, outpack, prev);
        const int c0 = CB * (s - u), cl = c0 + CB - 1;
        if (u <= ulast && cl >= cs && c0 <= ce) {
            // All lanes cut their next 32 columns out of the bitmap at the same step, every 16th (a lane on its own
            // would do it whenever ITS window runs out: with the lanes two columns apart that is one lane every step, and
            // the whole warp waits for that lane's loads); a lane that has just taken a new chunk (32 columns back) or
            // entered the band between two such steps fetches for itself.
            int j = c0 - tbase;
            if ((s & 15) == 0 || (unsigned)j > 30u) { bb_fetch_target_planes(P, c0, tb0, tb1, tok); tbase = c0; j = 0; }
            const uint32_t k0 = tb0 >> j, k1 = tb1 >> j, kk = tok >> j;  // bits 0, 1: columns c0, c0 + 1
            const int hin0 = c0 <= ce_up ? (int)((in >> 22) & 3u) - 1 : 1;
            const int hin1 = cl <= ce_up ? (int)((in >> 24) & 3u) - 1 : 1;
            int o0, o1, last_score;
            // With a and b even (bb_task_band makes them so) a chunk's columns are whole pairs - cs is even, ce odd unless
            // it is the pass's last column - and taking or leaving a chunk does not leave the common path: the lanes that
            // do either run a few extra instructions around the same two Myers steps as everybody else.  (Handled column
            // by column, as any odd band still is, every entry and exit made the whole warp run the general path for one
            // lane: a third of the warp instructions of a narrow pass.)
            const bool entered = even_band && c0 == cs;
            if (entered) enter_chunk(hin0, (int)(in & BB_MAX_SCORE));
            if ((even_band ? c0 >= cs : c0 > cs) && cl < ce + (even_band ? 1 : 0) && cl < ncols - 1 && (kk & 3u) == 3u && qok) {
                uint32_t Eq[L];
                planes(k0, k1, Eq);
                o0 = myers(Eq, hin0);
                planes(k0 >> 1, k1 >> 1, Eq);
                o1 = myers(Eq, hin1);
                score += o0 + o1;
                last_score = score;
                if (cl == ce) {  // (even band) the band has moved past this chunk: chunk u + K is next
                    u += K;
                    cs = max(0, CH * u - b);
                    ce = min(ncols - 1, CH * u + CH - 1 + a);
                    ce_up = min(ncols - 1, CH * u - 1 + a);
                }
            } else {
                int above = (int)(in & BB_MAX_SCORE);  // the chunk above after column c0 + h
                bool moved = false;
                o0 = 0; o1 = 0;
                last_score = score;
                if (c0 >= cs && c0 <= ce) {
                    o0 = column(c0, k0, k1, (kk & 1u) != 0u, hin0, above, entered);
                    last_score = score;
                    if (c0 == ce) moved = true;
                }
                above += (int)((in >> 24) & 3u) - 1;
                if (!moved && cl >= cs && cl <= ce) {
                    o1 = column(cl, k0 >> 1, k1 >> 1, (kk & 2u) != 0u, hin1, above, false);
                    last_score = score;
                    if (cl == ce) moved = true;
                } else o1 = 0;
                if (moved) {  // the band has moved past this chunk: chunk u + K is next
                    u += K;
                    cs = max(0, CH * u - b);
                    ce = min(ncols - 1, CH * u + CH - 1 + a);
                    ce_up = min(ncols - 1, CH * u - 1 + a);
                }
            }
            // the score after column c0 and the two horizontal outputs
            outpack = ((uint32_t)(last_score - o1) & BB_MAX_SCORE) | ((uint32_t)(o0 + 1) << 22) | ((uint32_t)(o1 + 1) << 24);
        }
    }
    const int owner = (lane & ~(K - 1)) | ((n > 0 ? (n - 1) / CH : 0) & (K - 1));
    result = __shfl_sync(BB_FULL, result, owner);
    __syncwarp();
    return result;
}

// Smallest L in {1,2,4,...,MAXL} with (a + b) / (32 L) + 2 <= K, or 0 if none.  MAXL bounds the variants a kernel
// instantiates (and with them its register footprint).
template <int MAXL>
__device__ __forceinline__ int bb_pick_L(int a, int b, int K) {
    for (int L = 1; L <= MAXL; L <<= 1)
        if ((a + b) / (32 * L) + 2 <= K) return L;
    return 0;
}

template <bool HIST, bool COLS, int MAXL, bool SM = false>
__device__ __forceinline__ int bb_band_dispatch(const BBProb &P, int K, int L) {
    if (MAXL >= 32 && L == 32) return bb_band_pass<(MAXL >= 32 ? 32 : 1), HIST, COLS, SM>(P, K);
    if (MAXL >= 16 && L == 16) return bb_band_pass<(MAXL >= 16 ? 16 : 1), HIST, COLS, SM>(P, K);
    if (MAXL >= 8 && L == 8) return bb_band_pass<(MAXL >= 8 ? 8 : 1), HIST, COLS, SM>(P, K);
    if (MAXL >= 4 && L == 4) return bb_band_pass<(MAXL >= 4 ? 4 : 1), HIST, COLS>(P, K);
    if (MAXL >= 2 && L == 2) return bb_band_pass<(MAXL >= 2 ? 2 : 1), HIST, COLS>(P, K);
    return bb_band_pass<1, HIST, COLS>(P, K);
}

// Fallback for bands wider than 32*32*30 rows: 1024-row strips, lane l owns one word of the strip and works on
// column (step - l); carries between strips go through hbuf.  Same outputs as bb_band_pass.
template <bool HIST, bool COLS>
__device__ int bb_strip_pass(const uint8_t *q, int qs, int n, const uint8_t *t, int ts, int ncols, int a, int b,
                             uint2 *hist, int nb_alloc, int *cols_out, int cols_lo, int8_t *hbuf) {
    const int lane = threadIdx.x & 31;
    const int nblk = (n + 31) >> 5;
    const int nstrips = (n + 1023) >> 10;
    int result = BB_INF;
    int bprev = 0;
    const int cols_hi = min(n - 1, ncols - 1 + b);
    for (int s = 0; s < nstrips; s++) {
        const int jstart = max(0, 1024 * s - b);
        const int jend = min(ncols - 1, 1024 * s + 1023 + a);
        if (jstart > jend) break;
        const int jstart_next = max(0, 1024 * (s + 1) - b);
        const int jend_prev = s > 0 ? min(ncols - 1, 1024 * (s - 1) + 1023 + a) : -1;
        const bool more_strips = (s + 1 < nstrips);
        const int row0 = 1024 * s + 32 * lane;
        uint32_t pA = 0, pC = 0, pG = 0, pT = 0, pO = 0;
        for (int r = 0; r < 32; r++) {
            const int row = row0 + r;
            if (row < n) {
                const uint8_t c = q[(long long)row * qs];
                const uint32_t bit = 1u << r;
                if (c == 'A') pA |= bit; else if (c == 'C') pC |= bit; else if (c == 'G') pG |= bit;
                else if (c == 'T') pT |= bit; else pO |= bit;
            }
        }
        uint32_t Pv = ~0u, Mv = 0u;
        int score = (jstart == 0 ? 1024 * s : bprev) + 32 * (lane + 1);
        int next_rec = BB_INF;
        uint32_t outpack = 0, prepack = 0;
        const int nsteps = (jend - jstart + 1) + 31;
        const int blk = 32 * s + lane;
        __syncwarp();
        for (int tau = 0; tau < nsteps; tau++) {
            if ((tau & 31) == 0) {
                const int col = jstart + tau + lane;
                uint32_t tc = 0; int h = 1;
                if (col <= jend) {
                    tc = t[(long long)col * ts];
                    if (s > 0 && col <= jend_prev) h = hbuf[col];
                }
                prepack = tc | ((uint32_t)(h + 1) << 8);
            }
            const uint32_t in0 = __shfl_sync(BB_FULL, prepack, tau & 31);
            const uint32_t inup = __shfl_up_sync(BB_FULL, outpack, 1);
            const uint32_t in = lane == 0 ? in0 : inup;
            const int col = jstart + tau - lane;
            const bool active = (tau >= lane) && (col <= jend);
            if (active) {
                const uint32_t tc = in & 0xffu;
                const int hin = (int)((in >> 8) & 3u) - 1;
                uint32_t Eq;
                if (tc == 'A') Eq = pA; else if (tc == 'C') Eq = pC; else if (tc == 'G') Eq = pG;
                else if (tc == 'T') Eq = pT;
                else {
                    Eq = 0;
                    uint32_t rest = pO;
                    while (rest) {
                        const int r = __ffs(rest) - 1;
                        rest &= rest - 1;
                        if (q[(long long)(row0 + r) * qs] == tc) Eq |= 1u << r;
                    }
                }
                const uint32_t hin_neg = hin < 0 ? 1u : 0u;
                const uint32_t Xv = Eq | Mv;
                Eq |= hin_neg;
                const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                uint32_t Ph = Mv | ~(Xh | Pv);
                uint32_t Mh = Pv & Xh;
                const int hout = (int)(Ph >> 31) - (int)(Mh >> 31);
                const uint32_t ph_raw = Ph;
                Ph = (Ph << 1) | (hin > 0 ? 1u : 0u);
                Mh = (Mh << 1) | hin_neg;
                Pv = Mh | ~(Xv | Ph);
                Mv = Ph & Xv;
                score += hout;
                outpack = tc | ((uint32_t)(hout + 1) << 8);
                if (HIST) {
                    const int bf = bb_first_block(col, a, nblk);
                    const int rel = blk - bf;
                    if (rel >= 0 && rel < nb_alloc && blk <= bb_last_block(col, b, n))
                        hist[col * nb_alloc + rel] = make_uint2(Pv, ph_raw);
                }
                if (lane == 31) {
                    if (more_strips) hbuf[col] = (int8_t)hout;
                    if (col == jstart_next - 1) next_rec = score;
                }
                if (col == ncols - 1) {
                    if (COLS) {
                        int run = score;
                        for (int r = 31; r >= 0; r--) {
                            const int row = row0 + r;
                            if (row < n && row >= cols_lo && row <= cols_hi) cols_out[row - cols_lo] = run;
                            run -= (int)((Pv >> r) & 1u) - (int)((Mv >> r) & 1u);
                        }
                    }
                    if (row0 <= n - 1 && n - 1 < row0 + 32) {
                        const int bit = (n - 1) - row0;
                        const uint32_t up = bit == 31 ? 0u : (Pv >> (bit + 1));
                        const uint32_t um = bit == 31 ? 0u : (Mv >> (bit + 1));
                        result = score - __popc(up) + __popc(um);
                    }
                }
            }
        }
        bprev = __shfl_sync(BB_FULL, next_rec, 31);
        __syncwarp();
    }
    const int owner = ((n - 1) & 1023) >> 5;
    result = __shfl_sync(BB_FULL, result, owner);
    return result;
}

// Traceback over the stored history of a leaf problem (edlib.cpp obtainAlignmentTraceback on exact deltas):
// at (i,j): 'I' if D[i][j]-D[i-1][j]==1, else 'D' if D[i][j]-D[i][j-1]==1, else '=' / 'X'.
// Each step inspects the 32 cells of the current diagonal, consumes the whole run of diagonal moves at once and
// then the single 'I'/'D' that ends it.  qbase: index of q[0] in the read (for emission).
template <bool EMIT>
__device__ void bb_traceback(const uint8_t *q, int n, const uint8_t *t, int m, int a, int b, const uint2 *hist,
                             int nb_alloc, BBEmit em, int qbase, BBAlnCounts &cnt) {
    const int lane = threadIdx.x & 31;
    const int nblk = (n + 31) >> 5;
    int i = n - 1, j = m - 1;
    int matches = 0, dels = 0;
    while (i >= 0 && j >= 0) {
        const int ii = i - lane, jj = j - lane;
        const bool valid = ii >= 0 && jj >= 0;
        bool up = false, left = false, diag = false, eq = false;
        if (valid) {
            const int bf = bb_first_block(jj, a, nblk);
            const int blk = ii >> 5;
            const int rel = blk - bf;
            if (rel >= 0 && rel < nb_alloc && blk <= bb_last_block(jj, b, n)) {
                const uint2 e = hist[jj * nb_alloc + rel];
                const int bit = ii & 31;
                up = (e.x >> bit) & 1u;
                left = !up && ((e.y >> bit) & 1u);
                diag = !up && !left;
            }
            eq = q[ii] == t[jj];
        }
        const uint32_t dmask = __ballot_sync(BB_FULL, diag);
        const uint32_t emask = __ballot_sync(BB_FULL, eq);
        const uint32_t umask = __ballot_sync(BB_FULL, up);
        const uint32_t lmask = __ballot_sync(BB_FULL, left);
        const int r = dmask == BB_FULL ? 32 : __ffs(~dmask) - 1;
        const uint32_t low = r == 32 ? BB_FULL : ((1u << r) - 1u);
        matches += __popc(emask & low);
        if (EMIT && lane < r) em.ops[qbase + ii] = eq ? BB_OP_EQ : BB_OP_X;
        i -= r; j -= r;
        if (r < 32 && i >= 0 && j >= 0) {
            if ((umask >> r) & 1u) {
                if (EMIT && lane == 0) em.ops[qbase + i] = BB_OP_I;
                i--;
            } else if ((lmask >> r) & 1u) {
                if (EMIT && lane == 0) atomicAdd(&em.dcnt[qbase + i], 1u);
                dels++;
                j--;
            } else {
                cnt.err |= 1;  // the canonical path left the stored band: cannot happen for a valid band
                break;
            }
        }
    }
    if (cnt.err == 0) {
        if (i >= 0) {  // column boundary reached: the remaining query characters are insertions
            if (EMIT) for (int x = lane; x <= i; x += 32) em.ops[qbase + x] = BB_OP_I;
        }
        if (j >= 0) {  // row boundary reached: the remaining target characters are deletions before q[0]
            dels += j + 1;
            if (EMIT && lane == 0) {
                if (qbase > 0) atomicAdd(&em.dcnt[qbase - 1], (unsigned int)(j + 1));
                else atomicAdd(em.lead_del, j + 1);
            }
        }
    }
    cnt.matches += matches;
    cnt.dels += dels;
}

// A leaf of edlib's recursion: forward pass with history, then traceback. k bounds the edit distance.
// q points at the leaf's first query character; qrel is that character's index in the emission arrays, qpeq its
// index in the read whose match bitmap is sc.peq.
template <bool EMIT, int MAXL>
__device__ int bb_leaf(const uint8_t *q, int n, const uint8_t *t, int m, int k, const BBScratch &sc, BBEmit em,
                       int qrel, int qpeq, BBAlnCounts &cnt) {
    int a, b;
    bb_band(n, m, k, a, b);
    const int nblk = (n + 31) >> 5;
    int nb_alloc = ((a + b) >> 5) + 2;
    if (nb_alloc > nblk) nb_alloc = nblk;
    if ((long long)nb_alloc * m > sc.hist_cap) { cnt.err |= 2; return BB_INF; }
    int d;
    const int L = bb_pick_L<MAXL>(a, b, 32);
    if (L > 0) {
        BBProb P;
        P.q = q; P.qs = 1; P.n = n; P.t = t; P.ts = 1; P.ncols = m; P.a = a; P.b = b;
        P.peq = sc.peq; P.peq_bit0 = qpeq + BB_PEQ_BIT0;
        P.hist = sc.hist; P.nb_alloc = nb_alloc; P.cols_out = nullptr; P.cols_lo = 0;
        d = bb_band_dispatch<true, false, MAXL>(P, 32, L);
    } else {
        if (m > sc.hbuf_cap) { cnt.err |= 2; return BB_INF; }
        d = bb_strip_pass<true, false>(q, 1, n, t, 1, m, a, b, sc.hist, nb_alloc, nullptr, 0, sc.hbuf);
    }
    __syncwarp();
    bb_traceback<EMIT>(q, n, t, m, a, b, sc.hist, nb_alloc, em, qrel, cnt);
    __syncwarp();
    return d;
}

__device__ __forceinline__ void bb_emit_all_deleted(int m, BBEmit em, int qbase, BBAlnCounts &cnt, bool emit) {
    // empty query: edlib.cpp obtainAlignment emits |t| deletions
    cnt.dels += m;
    if (emit && (threadIdx.x & 31) == 0) {
        if (qbase > 0) atomicAdd(&em.dcnt[qbase - 1], (unsigned int)m);
        else atomicAdd(em.lead_del, m);
    }
}

// edlib.cpp obtainAlignmentHirschberg's choice of the split row from the column scores in sc.L / sc.R, by one
// warp.  best < 0 (root): every optimal path crosses the split column, so the minimum sum is the edit distance.
static __device__ int bb_split_warp(const BBScratch &sc, int loL, int hiL, int loR, int hiR, int nn, int left_w, int right_w,
                             int &best, int &split, int &ls, int &rs) {
    const int lane = threadIdx.x & 31;
    int rlo = max(loL, nn - 2 - hiR); if (rlo < 0) rlo = 0;
    int rhi = min(hiL, nn - 2 - loR); if (rhi > nn - 2) rhi = nn - 2;
    const bool have_top = nn - 1 >= loR && nn - 1 <= hiR;  // empty query prefix on the left
    const bool have_bot = nn - 1 >= loL && nn - 1 <= hiL;  // empty query suffix on the right
    if (best < 0) {  // root: every optimal path crosses the split column, so the minimum sum is the distance
        int mn = BB_INF;
        for (int r = rlo + lane; r <= rhi; r += 32) mn = min(mn, sc.L[r - loL] + sc.R[(nn - 2 - r) - loR]);
        if (have_top) mn = min(mn, left_w + sc.R[(nn - 1) - loR]);
        if (have_bot) mn = min(mn, sc.L[(nn - 1) - loL] + right_w);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mn = min(mn, __shfl_xor_sync(BB_FULL, mn, d));
        best = mn;
    }
    // smallest interior row r in [0, nn-2] with L[r] + R[nn-2-r] == best
    split = -2; ls = 0; rs = 0;
    for (int base = rlo; base <= rhi; base += 32) {
        const int r = base + lane;
        bool hit = false;
        if (r <= rhi) hit = (sc.L[r - loL] + sc.R[(nn - 2 - r) - loR] == best);
        const uint32_t hm = __ballot_sync(BB_FULL, hit);
        if (hm) { split = base + __ffs(hm) - 1; break; }
    }
    if (split >= 0) { ls = sc.L[split - loL]; rs = sc.R[(nn - 2 - split) - loR]; }
    if (split == -2 && have_top) {
        const int v = sc.R[(nn - 1) - loR];
        if (left_w + v == best) { split = -1; ls = left_w; rs = v; }
    }
    if (split == -2 && have_bot) {
        const int v = sc.L[(nn - 1) - loL];
        if (v + right_w == best) { split = nn - 1; ls = v; rs = right_w; }
    }
    __syncwarp();
    return split == -2 ? 32 : 0;
}

// One Hirschberg node by one warp (edlib.cpp obtainAlignmentHirschberg): forward pass over the left half of the
// target, reverse pass over the right half, split row by edlib's rule.  q / t point at the read's first query /
// target character, the node is q[q0, q0+nn) x t[t0, t0+mm); band (a, b) must admit every optimal path.
// best < 0 on entry (root): the minimum of forward + reverse scores over the split column is the edit distance and
// is returned in best.  Returns 0 or an error code.
template <int MAXL>
__device__ int bb_node_warp(const uint8_t *q, const uint8_t *t, int q0, int nn, int t0, int mm, int a, int b,
                            const BBScratch &sc, int &best, int &split, int &ls, int &rs, int qabs = 0) {
    const int lane = threadIdx.x & 31;
    const int left_w = mm / 2, right_w = mm - left_w;
    const int loL = max(0, left_w - 1 - a), hiL = min(nn - 1, left_w - 1 + b);
    const int loR = max(0, right_w - 1 - a), hiR = min(nn - 1, right_w - 1 + b);
    if (hiL - loL + 1 > sc.lr_cap || hiR - loR + 1 > sc.lr_cap) return 16;
    {
        auto make_prob = [&](bool rev) {
            BBProb P;
            P.n = nn; P.a = a; P.b = b; P.peq = sc.peq; P.hist = nullptr; P.nb_alloc = 0;
            if (!rev) {
                P.q = q + q0; P.qs = 1; P.t = t + t0; P.ts = 1; P.ncols = left_w;
                P.peq_bit0 = qabs + q0 + BB_PEQ_BIT0; P.cols_out = sc.L; P.cols_lo = loL;
            } else {
                P.q = q + q0 + nn - 1; P.qs = -1; P.t = t + t0 + mm - 1; P.ts = -1; P.ncols = right_w;
                P.peq_bit0 = qabs + q0 + nn - 1 + BB_PEQ_BIT0; P.cols_out = sc.R; P.cols_lo = loR;
            }
            return P;
        };
        const int L2 = bb_pick_L<MAXL>(a, b, 16);
        if (L2 > 0) {  // forward and reverse pass side by side in two 16-lane groups
            const BBProb PG = make_prob(lane >= 16);
            bb_band_dispatch<false, true, MAXL>(PG, 16, L2);
        } else {
            const int L1 = bb_pick_L<MAXL>(a, b, 32);
            if (L1 > 0) {
                bb_band_dispatch<false, true, MAXL>(make_prob(false), 32, L1);
                bb_band_dispatch<false, true, MAXL>(make_prob(true), 32, L1);
            } else {
                if (mm > sc.hbuf_cap) return 4;
                bb_strip_pass<false, true>(q + q0, 1, nn, t + t0, 1, left_w, a, b, nullptr, 0, sc.L, loL, sc.hbuf);
                bb_strip_pass<false, true>(q + q0 + nn - 1, -1, nn, t + t0 + mm - 1, -1, right_w, a, b, nullptr, 0,
                                           sc.R, loR, sc.hbuf);
            }
        }
    }
    __syncwarp();
    return bb_split_warp(sc, loL, hiL, loR, hiR, nn, left_w, right_w, best, split, ls, rs);
}

// edlib.align(q, t, task='path') for one pair by one warp. q[0] is character `qabs` of the read whose match
// bitmap is sc.peq (built with bb_build_peq). k_upper >= edit distance (the caller knows how many edits it
// injected). Results accumulate into cnt; with EMIT the per-base ops / deletion counts are written at indices
// relative to q[0].
template <bool EMIT, int MAXL>
__device__ void bb_align(const uint8_t *q, int n, const uint8_t *t, int m, int k_upper, const BBScratch &sc,
                         BBEmit em, int qabs, BBAlnCounts &cnt) {
    const int lane = threadIdx.x & 31;
    {
        const int diff = n > m ? n - m : m - n;
        if (k_upper < diff) k_upper = diff;
        const int mx = n > m ? n : m;
        if (k_upper > mx) k_upper = mx;
    }
    if (n + m >= BB_MAX_SCORE) { cnt.err |= 128; return; }
    if (bb_uses_traceback(n, m)) {
        cnt.dist = bb_leaf<EMIT, MAXL>(q, n, t, m, k_upper, sc, em, 0, qabs, cnt);
        return;
    }
    int a, b;
    bb_band(n, m, k_upper, a, b);
    int best_root;
    {
        const int L = bb_pick_L<MAXL>(a, b, 32);
        if (L > 0) {
            BBProb P;
            P.q = q; P.qs = 1; P.n = n; P.t = t; P.ts = 1; P.ncols = m; P.a = a; P.b = b;
            P.peq = sc.peq; P.peq_bit0 = qabs + BB_PEQ_BIT0;
            P.hist = nullptr; P.nb_alloc = 0; P.cols_out = nullptr; P.cols_lo = 0;
            best_root = bb_band_dispatch<false, false, MAXL>(P, 32, L);
        } else {
            if (m > sc.hbuf_cap) { cnt.err |= 4; return; }
            best_root = bb_strip_pass<false, false>(q, 1, n, t, 1, m, a, b, nullptr, 0, nullptr, 0, sc.hbuf);
        }
    }
    cnt.dist = best_root;
    // depth-first Hirschberg (edlib.cpp obtainAlignmentHirschberg); left child is processed first so that
    // deletions in front of a leaf are credited to the query base that precedes them
    int sp = 0;
    int *stk = sc.stack;
    if (lane == 0) { stk[0] = 0; stk[1] = n; stk[2] = 0; stk[3] = m; stk[4] = best_root; }
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        sp--;
        const int q0 = stk[sp * 5 + 0], nn = stk[sp * 5 + 1], t0 = stk[sp * 5 + 2], mm = stk[sp * 5 + 3];
        const int best = stk[sp * 5 + 4];
        __syncwarp();
        if (nn == 0) { bb_emit_all_deleted(mm, em, q0, cnt, EMIT); continue; }
        if (mm == 0) {
            if (EMIT) for (int x = lane; x < nn; x += 32) em.ops[q0 + x] = BB_OP_I;
            continue;
        }
        if (bb_uses_traceback(nn, mm)) {
            const int d = bb_leaf<EMIT, MAXL>(q + q0, nn, t + t0, mm, best, sc, em, q0, qabs + q0, cnt);
            if (d != best) {
                cnt.err |= 8;
#ifdef BB_EMU_DEBUG
                if (lane == 0) printf("leaf mismatch: q0=%d nn=%d t0=%d mm=%d best=%d d=%d\n", q0, nn, t0, mm, best, d);
#endif
            }
            if (cnt.err) return;
            continue;
        }
        bb_band(nn, mm, best, a, b);
        const int left_w = mm / 2;
        int split = 0, ls = 0, rs = 0, node_best = best;
        const int nerr = bb_node_warp<MAXL>(q, t, q0, nn, t0, mm, a, b, sc, node_best, split, ls, rs, qabs);
        if (nerr) { cnt.err |= nerr; return; }
        const int right_w = mm - left_w;
        if (sp + 2 > sc.stack_cap) { cnt.err |= 64; return; }
        __syncwarp();
        if (lane == 0) {
            int *p = stk + sp * 5;  // right child below, left child on top
            p[0] = q0 + split + 1; p[1] = nn - split - 1; p[2] = t0 + left_w; p[3] = right_w; p[4] = rs;
            p[5] = q0; p[6] = split + 1; p[7] = t0; p[8] = left_w; p[9] = ls;
        }
        sp += 2;
        __syncwarp();
    }
}
