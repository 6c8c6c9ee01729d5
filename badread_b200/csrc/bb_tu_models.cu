// bb_tu_models.cu — the counting passes of the model builders on the GPU:
//   bb_count_kmer_alternatives   badread error_model   (error_model.py:31-83: which read k-mers each reference k-mer became)
//   bb_count_cigar_qscores       badread qscore_model  (qscore_model.py:78-153: quality of the middle base per CIGAR window)
// The reference walks every alignment column by column in Python, rebuilding a window string per step; here an
// alignment is a CTA, a window is a thread and a (window content) is a 64-bit key in an open-addressing table:
// count, first occurrence (the reference's dicts keep insertion order, and its stable sorts break ties by it) and,
// for the qscore model, a histogram of the 94 quality values.  Windows whose content does not fit a key (read k-mers /
// CIGARs longer than the key holds: a handful per million) go to an overflow list that the host evaluates exactly.
// Input per alignment a (host code: badread_b200/model_builders.py): the aligned slice of the read (and its
// qualities), the aligned slice of the reference already on the read's strand, and the CIGAR runs in read orientation
// with the read / reference offset each run starts at.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/badread_b200.h"

#define BBM_EMPTY 0xffffffffffffffffull
#define BBM_NQ 94   // quality characters '!' .. '~'

namespace {

struct Aln {
    const uint8_t *read, *qual, *ref;
    const int64_t *read_off, *ref_off, *ops_off;
    const uint32_t *ops;       // (len << 2) | type, type 0 = M, 1 = I, 2 = D
    const int32_t *op_read0;   // read offset (within the alignment) at which the run starts
    const int32_t *op_ref0;    // reference offset at which the run starts
};

struct Table {
    unsigned long long *keys, *first;
    unsigned int *counts;      // error model: one per slot; qscore model: BBM_NQ per slot
    long long cap;             // power of two
    int *status;               // [0]: table full, [1]: overflow list full
    int *ovf_aln, *ovf_pos, *ovf_k;
    unsigned long long *n_ovf;
    long long ovf_cap;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// Slot of `key` (claimed if new), or -1 if the table is (as good as) full: a probe sequence of 1024 slots does not
// happen below a load of ~0.95, and once one thread has given up the others stop at their next key instead of walking the
// whole table each - the caller doubles the table and counts again.
__device__ long long table_slot(const Table &T, unsigned long long key) {
    if (*(volatile int *)&T.status[0]) return -1;
    unsigned long long h = mix64(key) & (unsigned long long)(T.cap - 1);
    const long long limit = T.cap < 1024 ? T.cap : 1024;
    for (long long probe = 0; probe < limit; probe++) {
        const unsigned long long prev = atomicCAS(&T.keys[h], BBM_EMPTY, key);
        if (prev == BBM_EMPTY || prev == key) return (long long)h;
        h = (h + 1) & (unsigned long long)(T.cap - 1);
    }
    atomicExch(&T.status[0], 1);
    return -1;
}

__device__ void overflow(const Table &T, int aln, int pos, int k) {
    const unsigned long long i = atomicAdd(T.n_ovf, 1ull);
    if ((long long)i < T.ovf_cap) { T.ovf_aln[i] = aln; T.ovf_pos[i] = pos; T.ovf_k[i] = k; }
    else atomicExch(&T.status[1], 1);
}

__device__ __forceinline__ int base_code(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// ---------------------------------------------------------------------------------------------- error model
// error_model.py:45-66.  The window of reference base r (r = 0 .. n_ref - k) spans the alignment columns from r's column
// to the column of reference base r + k - 1; its read k-mer is read[rp[r] : rp[r+k-1] + isM[r+k-1]) with rp[x] = read
// bases in front of x's column (the first window starts at column 0, i.e. at read base 0, whatever the alignment starts
// with).  Counted if the read k-mer has more than one base, both k-mers are ACGT only and they agree in their first and
// last base.  Key: reference k-mer (2k bits) | length (6 bits) | read k-mer (2 bits a base).
__global__ void __launch_bounds__(256) k_kmer_alternatives(Aln A, int n_aln, int k, int *rp_pool, uint8_t *ism_pool, Table T) {
    const int a = blockIdx.x;
    if (a >= n_aln) return;
    const uint8_t *read = A.read + A.read_off[a], *ref = A.ref + A.ref_off[a];
    const int n_ref = (int)(A.ref_off[a + 1] - A.ref_off[a]);
    int *rp = rp_pool + A.ref_off[a];
    uint8_t *ism = ism_pool + A.ref_off[a];
    for (long long o = A.ops_off[a] + threadIdx.x; o < A.ops_off[a + 1]; o += blockDim.x) {
        const uint32_t op = A.ops[o];
        const int len = (int)(op >> 2), type = (int)(op & 3u), p0 = A.op_read0[o], r0 = A.op_ref0[o];
        if (type == 0) for (int i = 0; i < len; i++) { rp[r0 + i] = p0 + i; ism[r0 + i] = 1; }
        else if (type == 2) for (int i = 0; i < len; i++) { rp[r0 + i] = p0; ism[r0 + i] = 0; }
    }
    __syncthreads();
    const int shift_ref = 64 - 2 * k, shift_len = shift_ref - 6, max_len = shift_len / 2;
    for (int r = threadIdx.x; r + k <= n_ref; r += blockDim.x) {
        const int p_lo = r == 0 ? 0 : rp[r], p_hi = rp[r + k - 1] + ism[r + k - 1];
        const int len = p_hi - p_lo;
        if (len <= 1) continue;
        if (read[p_lo] != ref[r] || read[p_hi - 1] != ref[r + k - 1]) continue;
        unsigned long long key = 0;
        bool ok = true;
        for (int j = 0; j < k; j++) {
            const int c = base_code(ref[r + j]);
            ok = ok && c >= 0;
            key = (key << 2) | (unsigned long long)(c & 3);
        }
        if (!ok) continue;
        key <<= shift_ref;
        if (len > max_len) {  // (the host checks the read k-mer's alphabet itself)
            overflow(T, a, r, k);
            continue;
        }
        unsigned long long rb = 0;
        for (int j = 0; j < len; j++) {
            const int c = base_code(read[p_lo + j]);
            ok = ok && c >= 0;
            rb |= (unsigned long long)(c & 3) << (2 * j);
        }
        if (!ok) continue;
        key |= ((unsigned long long)len << shift_len) | rb;
        const long long s = table_slot(T, key);
        if (s < 0) return;
        atomicAdd(&T.counts[s], 1u);
        atomicMin(&T.first[s], ((unsigned long long)a << 32) | (unsigned long long)r);
    }
}

// ---------------------------------------------------------------------------------------------- qscore model
// qscore_model.py:104-141.  Per read base i: sym[i] ('=' 0, 'X' 1, 'I' 2) and dcount[i] = 'D' columns between base i and
// base i + 1.  The window of (k, i) - k = 1, 3, ..., K read bases from base i on - has the CIGAR
// sym[i] D^min(dcount[i], max_del) sym[i+1] ... sym[i+k-1] (the first window of an alignment also takes the 'D' columns in
// front of base 0) and the quality of its middle base.  Key: length (6 bits) | 2 bits a symbol ('D' = 3).
__global__ void __launch_bounds__(256) k_cigar_qscores(Aln A, int n_aln, int K, int max_del, uint8_t *sym_pool, int *dc_pool,
                                                       int *lead_pool, Table T, unsigned long long *overall) {
    const int a = blockIdx.x;
    if (a >= n_aln) return;
    const uint8_t *read = A.read + A.read_off[a], *qual = A.qual + A.read_off[a], *ref = A.ref + A.ref_off[a];
    const int n_read = (int)(A.read_off[a + 1] - A.read_off[a]);
    uint8_t *sym = sym_pool + A.read_off[a];
    int *dc = dc_pool + A.read_off[a];
    for (int i = threadIdx.x; i < n_read; i += blockDim.x) dc[i] = 0;
    if (threadIdx.x == 0) lead_pool[a] = 0;
    __syncthreads();
    for (long long o = A.ops_off[a] + threadIdx.x; o < A.ops_off[a + 1]; o += blockDim.x) {
        const uint32_t op = A.ops[o];
        const int len = (int)(op >> 2), type = (int)(op & 3u), p0 = A.op_read0[o], r0 = A.op_ref0[o];
        if (type == 0) for (int i = 0; i < len; i++) sym[p0 + i] = read[p0 + i] == ref[r0 + i] ? 0 : 1;
        else if (type == 1) for (int i = 0; i < len; i++) sym[p0 + i] = 2;
        else if (p0 > 0) atomicAdd(&dc[p0 - 1], len);   // (two 'D' runs in a row are one run of 'D' columns)
        else atomicAdd(&lead_pool[a], len);
    }
    __syncthreads();
    const int lead = lead_pool[a];
    for (int kk = 1, kidx = 0; kk <= K; kk += 2, kidx++) {
        for (int i = threadIdx.x; i + kk <= n_read; i += blockDim.x) {
            unsigned long long key = 0;
            int len = 0;
            bool fits = true;
            auto push = [&](unsigned long long s, int count) {
                for (int x = 0; x < count; x++) {
                    if (len >= 29) { fits = false; return; }
                    key |= s << (2 * len);
                    len++;
                }
            };
            if (i == 0) push(3ull, lead < max_del ? lead : max_del);
            for (int j = 0; j < kk && fits; j++) {
                push((unsigned long long)sym[i + j], 1);
                if (j + 1 < kk) push(3ull, dc[i + j] < max_del ? dc[i + j] : max_del);
            }
            const int q = (int)qual[i + (kk - 1) / 2] - 33;
            if (q < 0 || q >= BBM_NQ) { overflow(T, a, i, -kk); continue; }   // not a quality character: the host decides
            if (kk == 1) atomicAdd(&overall[q], 1ull);
            if (!fits) { overflow(T, a, i, kk); continue; }
            key |= (unsigned long long)len << 58;
            const long long s = table_slot(T, key);
            if (s < 0) return;
            atomicAdd(&T.counts[s * BBM_NQ + q], 1u);
            atomicMin(&T.first[s], ((unsigned long long)a << 36) | ((unsigned long long)kidx << 32) | (unsigned long long)i);
        }
    }
}

// Occupied slots -> dense output (arbitrary order; the host sorts by first occurrence).
__global__ void k_compact(Table T, int per_slot, unsigned long long *keys_out, unsigned long long *first_out,
                          unsigned int *counts_out, unsigned long long *n_out, long long out_cap) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T.cap || T.keys[s] == BBM_EMPTY) return;
    const unsigned long long i = atomicAdd(n_out, 1ull);
    if ((long long)i >= out_cap) return;
    keys_out[i] = T.keys[s];
    first_out[i] = T.first[s];
    for (int x = 0; x < per_slot; x++) counts_out[i * per_slot + x] = T.counts[s * per_slot + x];
}

struct DevMem {   // everything a call allocates, released on every exit path
    void *p[32];
    int n = 0;
    ~DevMem() { for (int i = 0; i < n; i++) cudaFree(p[i]); }
    template <typename X>
    cudaError_t get(X **out, size_t bytes, const void *src = nullptr, int fill = -1) {
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
        if (e != cudaSuccess) return e;
        p[n++] = q;
        *out = (X *)q;
        if (src) e = cudaMemcpy(q, src, bytes, cudaMemcpyHostToDevice);
        else if (fill >= 0) e = cudaMemset(q, fill, bytes ? bytes : 16);
        return e;
    }
};

#define BBM_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    std::snprintf(g_model_error, sizeof(g_model_error), "%s: %s", #call, cudaGetErrorString(e_)); return BB_ERR_CUDA; } } while (0)

thread_local char g_model_error[256] = "";

int count_common(bool qscores, int device, int k, int max_del, int32_t n_aln, const uint8_t *read, const uint8_t *qual,
                 const int64_t *read_off, const uint8_t *ref, const int64_t *ref_off, const uint32_t *ops, const int32_t *op_read0,
                 const int32_t *op_ref0, const int64_t *ops_off, int64_t table_cap, uint64_t *keys_out, uint64_t *first_out,
                 uint32_t *counts_out, int64_t *n_entries, uint64_t *overall_out, int64_t ovf_cap, int32_t *ovf_aln,
                 int32_t *ovf_pos, int32_t *ovf_k, int64_t *n_ovf) {
    g_model_error[0] = 0;
    if (n_aln <= 0 || !read || !read_off || !ref || !ref_off || !ops || !ops_off || !op_read0 || !op_ref0 || !keys_out ||
        !first_out || !counts_out || !n_entries || !n_ovf || table_cap < 16 || (table_cap & (table_cap - 1)) ||
        (qscores && (!qual || !overall_out || k < 1 || k > 13 || !(k & 1) || max_del < 0)) || (!qscores && (k < 1 || k > 12))) {
        std::snprintf(g_model_error, sizeof(g_model_error), "bb_count_*: invalid argument");
        return BB_ERR_ARG;
    }
    BBM_TRY(cudaSetDevice(device));
    const int64_t n_read = read_off[n_aln], n_ref = ref_off[n_aln], n_ops = ops_off[n_aln];
    const int per_slot = qscores ? BBM_NQ : 1;
    DevMem mem;
    Aln A{};
    uint8_t *d_read, *d_qual = nullptr, *d_ref;
    int64_t *d_read_off, *d_ref_off, *d_ops_off;
    uint32_t *d_ops;
    int32_t *d_p0, *d_r0;
    BBM_TRY(mem.get(&d_read, (size_t)n_read, read));
    if (qscores) BBM_TRY(mem.get(&d_qual, (size_t)n_read, qual));
    BBM_TRY(mem.get(&d_ref, (size_t)n_ref, ref));
    BBM_TRY(mem.get(&d_read_off, (size_t)(n_aln + 1) * 8, read_off));
    BBM_TRY(mem.get(&d_ref_off, (size_t)(n_aln + 1) * 8, ref_off));
    BBM_TRY(mem.get(&d_ops_off, (size_t)(n_aln + 1) * 8, ops_off));
    BBM_TRY(mem.get(&d_ops, (size_t)n_ops * 4, ops));
    BBM_TRY(mem.get(&d_p0, (size_t)n_ops * 4, op_read0));
    BBM_TRY(mem.get(&d_r0, (size_t)n_ops * 4, op_ref0));
    A.read = d_read; A.qual = d_qual; A.ref = d_ref; A.read_off = d_read_off; A.ref_off = d_ref_off; A.ops_off = d_ops_off;
    A.ops = d_ops; A.op_read0 = d_p0; A.op_ref0 = d_r0;
    Table T{};
    T.cap = table_cap; T.ovf_cap = ovf_cap;
    BBM_TRY(mem.get(&T.keys, (size_t)table_cap * 8, nullptr, 0xff));
    BBM_TRY(mem.get(&T.first, (size_t)table_cap * 8, nullptr, 0xff));
    BBM_TRY(mem.get(&T.counts, (size_t)table_cap * per_slot * 4, nullptr, 0));
    BBM_TRY(mem.get(&T.status, 16, nullptr, 0));
    BBM_TRY(mem.get(&T.n_ovf, 16, nullptr, 0));
    BBM_TRY(mem.get(&T.ovf_aln, (size_t)ovf_cap * 4));
    BBM_TRY(mem.get(&T.ovf_pos, (size_t)ovf_cap * 4));
    BBM_TRY(mem.get(&T.ovf_k, (size_t)ovf_cap * 4));
    unsigned long long *d_overall = nullptr;
    if (qscores) {
        uint8_t *sym; int *dc, *lead;
        BBM_TRY(mem.get(&sym, (size_t)n_read));
        BBM_TRY(mem.get(&dc, (size_t)n_read * 4));
        BBM_TRY(mem.get(&lead, (size_t)n_aln * 4));
        BBM_TRY(mem.get(&d_overall, BBM_NQ * 8, nullptr, 0));
        k_cigar_qscores<<<n_aln, 256>>>(A, n_aln, k, max_del, sym, dc, lead, T, d_overall);
    } else {
        int *rp; uint8_t *ism;
        BBM_TRY(mem.get(&rp, (size_t)n_ref * 4));
        BBM_TRY(mem.get(&ism, (size_t)n_ref));
        k_kmer_alternatives<<<n_aln, 256>>>(A, n_aln, k, rp, ism, T);
    }
    BBM_TRY(cudaGetLastError());
    unsigned long long *d_keys_out, *d_first_out, *d_n;
    unsigned int *d_counts_out;
    BBM_TRY(mem.get(&d_keys_out, (size_t)table_cap * 8));
    BBM_TRY(mem.get(&d_first_out, (size_t)table_cap * 8));
    BBM_TRY(mem.get(&d_counts_out, (size_t)table_cap * per_slot * 4));
    BBM_TRY(mem.get(&d_n, 16, nullptr, 0));
    k_compact<<<(unsigned int)((table_cap + 255) / 256), 256>>>(T, per_slot, d_keys_out, d_first_out, d_counts_out, d_n, table_cap);
    BBM_TRY(cudaGetLastError());
    int status[2] = {0, 0};
    unsigned long long n = 0, novf = 0;
    BBM_TRY(cudaMemcpy(status, T.status, sizeof(status), cudaMemcpyDeviceToHost));   // (synchronizes with the kernels)
    BBM_TRY(cudaMemcpy(&n, d_n, 8, cudaMemcpyDeviceToHost));
    BBM_TRY(cudaMemcpy(&novf, T.n_ovf, 8, cudaMemcpyDeviceToHost));
    *n_entries = (int64_t)n; *n_ovf = (int64_t)novf;
    if (status[0] || status[1]) {
        std::snprintf(g_model_error, sizeof(g_model_error), "bb_count_*: %s too small", status[0] ? "table" : "overflow list");
        return BB_ERR_CAPACITY;
    }
    BBM_TRY(cudaMemcpy(keys_out, d_keys_out, (size_t)n * 8, cudaMemcpyDeviceToHost));
    BBM_TRY(cudaMemcpy(first_out, d_first_out, (size_t)n * 8, cudaMemcpyDeviceToHost));
    BBM_TRY(cudaMemcpy(counts_out, d_counts_out, (size_t)n * per_slot * 4, cudaMemcpyDeviceToHost));
    if (novf) {
        BBM_TRY(cudaMemcpy(ovf_aln, T.ovf_aln, (size_t)novf * 4, cudaMemcpyDeviceToHost));
        BBM_TRY(cudaMemcpy(ovf_pos, T.ovf_pos, (size_t)novf * 4, cudaMemcpyDeviceToHost));
        BBM_TRY(cudaMemcpy(ovf_k, T.ovf_k, (size_t)novf * 4, cudaMemcpyDeviceToHost));
    }
    if (qscores) BBM_TRY(cudaMemcpy(overall_out, d_overall, BBM_NQ * 8, cudaMemcpyDeviceToHost));
    return BB_OK;
}

}  // namespace

extern "C" const char *bb_model_error(void) { return g_model_error; }

extern "C" int bb_count_kmer_alternatives(int device, int k, int32_t n_aln, const uint8_t *read, const int64_t *read_off,
                                          const uint8_t *ref, const int64_t *ref_off, const uint32_t *ops,
                                          const int32_t *op_read0, const int32_t *op_ref0, const int64_t *ops_off,
                                          int64_t table_cap, uint64_t *keys_out, uint64_t *first_out, uint32_t *counts_out,
                                          int64_t *n_entries, int64_t ovf_cap, int32_t *ovf_aln, int32_t *ovf_pos,
                                          int32_t *ovf_k, int64_t *n_ovf) {
    return count_common(false, device, k, 0, n_aln, read, nullptr, read_off, ref, ref_off, ops, op_read0, op_ref0, ops_off,
                        table_cap, keys_out, first_out, counts_out, n_entries, nullptr, ovf_cap, ovf_aln, ovf_pos, ovf_k, n_ovf);
}

extern "C" int bb_count_cigar_qscores(int device, int k, int max_del, int32_t n_aln, const uint8_t *read, const uint8_t *qual,
                                      const int64_t *read_off, const uint8_t *ref, const int64_t *ref_off, const uint32_t *ops,
                                      const int32_t *op_read0, const int32_t *op_ref0, const int64_t *ops_off,
                                      int64_t table_cap, uint64_t *keys_out, uint64_t *first_out, uint32_t *counts_out,
                                      int64_t *n_entries, uint64_t *overall_out, int64_t ovf_cap, int32_t *ovf_aln,
                                      int32_t *ovf_pos, int32_t *ovf_k, int64_t *n_ovf) {
    return count_common(true, device, k, max_del, n_aln, read, qual, read_off, ref, ref_off, ops, op_read0, op_ref0, ops_off,
                        table_cap, keys_out, first_out, counts_out, n_entries, overall_out, ovf_cap, ovf_aln, ovf_pos, ovf_k,
                        n_ovf);
}
