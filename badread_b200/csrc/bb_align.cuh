// bb_align.cuh — warp-level global alignment with edlib's path semantics.
//
// Replaces edlib.align(query, target, mode='NW', task='path') at the reference's call sites
// (simulate.py:330,340 — query = original window, target = mutated window; qscore_model.py:37 —
// query = mutated read, target = original fragment).  edlib (third-party, not vendored) defines the path as:
//   * traceback from the bottom-right cell preferring UP ('I', consumes a query char) over LEFT ('D') over the
//     diagonal, when its traceback state estimate 20*ceil(|q|/64)*|t| + 8*|t| is below 1 MiB;
//   * otherwise Hirschberg on the target (split at |t|/2, smallest interior query row whose left+right scores
//     equal the best score, then row -1, then row |q|-1) recursing with the same switch.
// One warp aligns one pair.  Rows are query characters; lane l owns a 32-row word of the current 1024-row
// strip and runs the Myers/Hyyrö bit-vector recurrence on column (step - l): a 32-lane wavefront whose
// horizontal carries travel by __shfl_up.  Vertical (+1) and horizontal (+1) delta words of in-band blocks are
// kept per column ("history") so the traceback needs two bits per cell; the traceback itself walks whole
// diagonal runs per step with warp ballots.
#pragma once
#include <cstdint>

#define BB_FULL 0xffffffffu
#define BB_INF 0x3fffffff
#define BB_OP_EQ 0
#define BB_OP_X 1
#define BB_OP_I 2

struct BBScratch {
    uint2 *hist;     // (Pv, PhRaw) per (column, block - first_block(column))
    int hist_cap;    // entries
    int8_t *hbuf;    // per-column horizontal delta leaving the bottom of the previous strip
    int hbuf_cap;
    int *L, *R;      // Hirschberg column scores (forward / reverse), indexed by row - row_lo
    int lr_cap;
    int *stack;      // DFS stack, 5 ints per node
    int stack_cap;   // nodes
};

struct BBEmit {      // where the final alignment is written (nullptr members => counts only)
    uint8_t *ops;    // per query base: BB_OP_EQ / BB_OP_X / BB_OP_I
    uint16_t *dcnt;  // per query base: 'D' columns between this base and the next (saturating)
    int *lead_del;   // 'D' columns before the first query base
};

struct BBAlnCounts {
    int matches;     // '=' columns
    int dels;        // 'D' columns (alignment columns = query length + dels)
    int dist;        // edit distance
    int err;         // non-zero: invariant violated
};

__device__ __forceinline__ void bb_band(int n, int m, int k, int &a, int &b) {
    // a path of cost <= k from (0,0) to (n,m) has at most (k-(n-m))/2 'D' and (k+(n-m))/2 'I' moves:
    // every cell (i,j) it can visit satisfies j - a <= i <= j + b
    a = (k - (n - m)) / 2; if (a < 0) a = 0;
    b = (k + (n - m)) / 2; if (b < 0) b = 0;
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

// Banded NW over columns [0, ncols) of t (stride ts) against the n rows of q (stride qs).
// HIST: store (Pv, PhRaw) of in-band blocks.  COLS: write D[row][ncols-1] for the in-band rows of the last
// column to cols_out[row - cols_lo].  Returns D[n-1][ncols-1] when the last strip reaches the last column,
// else BB_INF.  Values are exact for every cell on a path of cost <= the k the band was derived from and
// upper bounds elsewhere.
template <bool HIST, bool COLS>
__device__ int bb_myers_pass(const uint8_t *q, int qs, int n, const uint8_t *t, int ts, int ncols, int a, int b,
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
        int score = (jstart == 0 ? 1024 * s : bprev) + 32 * (lane + 1);  // D at this lane's bottom row, column jstart-1
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
                else {  // non-ACGT target character: exact byte equality against the non-ACGT rows
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
    // the lane that owns row n-1 holds the result
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
                if (EMIT && lane == 0) {
                    const uint16_t v = em.dcnt[qbase + i];
                    if (v != 0xffff) em.dcnt[qbase + i] = v + 1;
                }
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
                if (qbase > 0) {
                    const uint32_t v = (uint32_t)em.dcnt[qbase - 1] + (uint32_t)(j + 1);
                    em.dcnt[qbase - 1] = v > 0xffffu ? 0xffff : (uint16_t)v;
                } else {
                    *em.lead_del += j + 1;
                }
            }
        }
    }
    cnt.matches += matches;
    cnt.dels += dels;
}

// A leaf of edlib's recursion: forward pass with history, then traceback. k bounds the edit distance.
template <bool EMIT>
__device__ int bb_leaf(const uint8_t *q, int n, const uint8_t *t, int m, int k, const BBScratch &sc, BBEmit em,
                       int qbase, BBAlnCounts &cnt) {
    int a, b;
    bb_band(n, m, k, a, b);
    const int nblk = (n + 31) >> 5;
    int nb_alloc = ((a + b) >> 5) + 2;
    if (nb_alloc > nblk) nb_alloc = nblk;
    if ((long long)nb_alloc * m > sc.hist_cap || m > sc.hbuf_cap) { cnt.err |= 2; return BB_INF; }
    const int d = bb_myers_pass<true, false>(q, 1, n, t, 1, m, a, b, sc.hist, nb_alloc, nullptr, 0, sc.hbuf);
    __syncwarp();
    bb_traceback<EMIT>(q, n, t, m, a, b, sc.hist, nb_alloc, em, qbase, cnt);
    __syncwarp();
    return d;
}

__device__ __forceinline__ void bb_emit_all_deleted(int m, BBEmit em, int qbase, BBAlnCounts &cnt, bool emit) {
    // empty query: edlib.cpp obtainAlignment emits |t| deletions
    cnt.dels += m;
    if (emit && (threadIdx.x & 31) == 0) {
        if (qbase > 0) {
            const uint32_t v = (uint32_t)em.dcnt[qbase - 1] + (uint32_t)m;
            em.dcnt[qbase - 1] = v > 0xffffu ? 0xffff : (uint16_t)v;
        } else {
            *em.lead_del += m;
        }
    }
}

// edlib.align(q, t, task='path') for one pair by one warp. k_upper >= edit distance (the caller knows how many
// edits it injected). Results accumulate into cnt; with EMIT the per-base ops / deletion counts are written.
template <bool EMIT>
__device__ void bb_align(const uint8_t *q, int n, const uint8_t *t, int m, int k_upper, const BBScratch &sc,
                         BBEmit em, BBAlnCounts &cnt) {
    const int lane = threadIdx.x & 31;
    {
        const int diff = n > m ? n - m : m - n;
        if (k_upper < diff) k_upper = diff;
        const int mx = n > m ? n : m;
        if (k_upper > mx) k_upper = mx;
    }
    if (bb_uses_traceback(n, m)) {
        cnt.dist = bb_leaf<EMIT>(q, n, t, m, k_upper, sc, em, 0, cnt);
        return;
    }
    if (m > sc.hbuf_cap) { cnt.err |= 4; return; }
    int a, b;
    bb_band(n, m, k_upper, a, b);
    const int best_root = bb_myers_pass<false, false>(q, 1, n, t, 1, m, a, b, nullptr, 0, nullptr, 0, sc.hbuf);
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
            const int d = bb_leaf<EMIT>(q + q0, nn, t + t0, mm, best, sc, em, q0, cnt);
            if (d != best) cnt.err |= 8;
            if (cnt.err) return;
            continue;
        }
        const int left_w = mm / 2, right_w = mm - left_w;
        bb_band(nn, mm, best, a, b);
        const int loL = max(0, left_w - 1 - a), hiL = min(nn - 1, left_w - 1 + b);
        const int loR = max(0, right_w - 1 - a), hiR = min(nn - 1, right_w - 1 + b);
        if (hiL - loL + 1 > sc.lr_cap || hiR - loR + 1 > sc.lr_cap) { cnt.err |= 16; return; }
        bb_myers_pass<false, true>(q + q0, 1, nn, t + t0, 1, left_w, a, b, nullptr, 0, sc.L, loL, sc.hbuf);
        bb_myers_pass<false, true>(q + q0 + nn - 1, -1, nn, t + t0 + mm - 1, -1, right_w, a, b, nullptr, 0, sc.R,
                                   loR, sc.hbuf);
        __syncwarp();
        // smallest interior row r in [0, nn-2] with L[r] + R[nn-2-r] == best
        int split = -2, ls = 0, rs = 0;
        {
            int rlo = max(loL, nn - 2 - hiR); if (rlo < 0) rlo = 0;
            int rhi = min(hiL, nn - 2 - loR); if (rhi > nn - 2) rhi = nn - 2;
            for (int base = rlo; base <= rhi; base += 32) {
                const int r = base + lane;
                bool hit = false;
                if (r <= rhi) hit = (sc.L[r - loL] + sc.R[(nn - 2 - r) - loR] == best);
                const uint32_t hm = __ballot_sync(BB_FULL, hit);
                if (hm) { split = base + __ffs(hm) - 1; break; }
            }
            if (split >= 0) { ls = sc.L[split - loL]; rs = sc.R[(nn - 2 - split) - loR]; }
        }
        if (split == -2 && nn - 1 >= loR && nn - 1 <= hiR) {  // empty query prefix on the left
            const int v = sc.R[(nn - 1) - loR];
            if (left_w + v == best) { split = -1; ls = left_w; rs = v; }
        }
        if (split == -2 && nn - 1 >= loL && nn - 1 <= hiL) {  // empty query suffix on the right
            const int v = sc.L[(nn - 1) - loL];
            if (v + right_w == best) { split = nn - 1; ls = v; rs = right_w; }
        }
        if (split == -2) { cnt.err |= 32; return; }
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
