// bb_api.cu — C ABI of libbadread_b200.so (see include/badread_b200.h): context, one-time uploads, batch
// orchestration. All hot-path work is done by the kernels in bb_kernels.cuh; there is no CPU path here.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "bb_kernels.cuh"
#include "bb_launch.h"

namespace {

struct DevBuf {  // grow-only device allocation
    void *p = nullptr;
    size_t cap = 0;
    bool borrowed = false;  // p belongs to another context (the workers of a context share its reference and tables)
    void borrow(const DevBuf &o) { release(); p = o.p; cap = o.cap; borrowed = true; }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && !borrowed) return cudaSuccess;
        if (borrowed) { p = nullptr; cap = 0; borrowed = false; }
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

const char *kStageNames[BB_N_STAGES] = {"build_fragments", "error_loop", "scan", "join", "final_align", "qscores",
                                        "compact", "total"};

}  // namespace

struct bb_ctx {
    int device = 0;
    int sm_count = 0;
    uint64_t seed = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaStream_t side[2][3] = {};   // per alignment pipeline: the streams of the node classes that run next to the main one
    cudaEvent_t ev_side[2][3] = {}, ev_level[2] = {};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    int64_t launches = 0;

    // reference + models
    DevBuf ref; int64_t ref_len = 0;
    bool have_em = false, have_qm = false;
    BBErrorModelDev em{}; DevBuf em_k2r, em_rowoff, em_cum, em_flags, em_slots, em_pool, em_rowinfo;
    BBQScoreModelDev qm{}; DevBuf qm_hkeys, qm_hvals, qm_rowoff, qm_scores, qm_cum;

    // batch
    int n_reads = 0;
    bool uploaded = false, ran = false;
    std::vector<BBReadDev> h_reads;
    std::vector<int32_t> h_inlen;
    std::vector<double> h_target;
    int64_t frag_total = 0, out_total = 0;
    int64_t seq_cap = 0, out_cap = 0, speq_cap = 0;  // capacities of the per-batch buffers (from the fragment lengths)
    int64_t log_total = 0, wres_total = 0;
    int max_len = 0;
    // sizing knobs a retry raises (bb_k_scan / the node kernels flag what did not fit; see w_finish)
    double slack = 1.25;       // joined reads may be this much longer than their fragments in total
    int n_rounds = 3;          // mutate -> windows -> replay rounds enqueued without asking the device in between
    int n_levels = 0;          // Hirschberg levels enqueued (from the longest fragment)
    int extra_levels = 0;
    bool lr_worst = false;     // size the split-score scratch for the worst case instead of the expected edit count
    struct RunInfo { BBScanOut scan; int counters[256]; int qcount[2][32]; } *h_info = nullptr;  // pinned
    std::vector<BBReadDev> h_res;  // per-read records of the finished run
    bool finished = false;
    bool reran = false;        // w_finish had to run the batch again (copies enqueued before that are stale)
    DevBuf d_scan;
    void *nccl_comm = nullptr;   // ncclComm_t of this context's device (bb_comm_init_rank / bb_comm_init_all)
    DevBuf d_red;                // two int64: send, receive of bb_allreduce_bases
    cudaEvent_t ev_scan = nullptr;
    DevBuf d_read_index, d_seg_off, d_segs, d_lit, d_target, d_order, d_reads;
    int n_lane_reads = 0, n_long_reads = 0;
    std::vector<int> h_order;
    DevBuf d_kidx, d_frag, d_state, d_seq, d_ops, d_dcnt, d_qual, d_out_seq, d_out_qual, d_counter, d_fpeq, d_speq, d_fallback;
    int64_t fpeq_total = 0;

    // scratch
    int n_warps = 0;
    BBScratchPool pool{}, pool_lean{};  // pool_lean: split-score arrays of the single-warp node kernels (bands < 2048 rows)
    DevBuf s_hist, s_hbuf, s_lr, s_stack, s_tbuf, s_peq, s_ltbuf, s_leafhist, s_lr_lean, s_wckpt, s_lanehist;
    DevBuf d_ctime, d_chlog, d_wres, d_wtasks, d_wfallback, d_active;
    DevBuf p_q, p_t, p_ops, p_dcnt, p_out, p_qual;  // single-pair entry points (bb_align_path / bb_get_qscores): kept between calls
    int lane8_cols = 4096;  // routing limit of the lane node kernel (tuning knob)
    int pair_ctas = 1;   // CTAs per SM of the warp-pair node kernel (tuning knob)
    bool head_worker = true;  // worker 0 = the longest reads only (see bb_batch_upload)
    bool is_head = false;     // this worker holds the head batch of the current upload
    bool lpt_order = true;    // node queues below the roots walked from the end (longest nodes first)
    int ring_t = 4;           // columns per traceback tick of the 4-word window aligner (2, 4 or 8)
    bool lowmem = false;      // window / leaf aligners with checkpoints + shared-memory tiles instead of global history
    bool use_quad = false;    // wide nodes by 8-warp CTAs (bb_k_node_quad) instead of warp pairs
    int grid_div_env = 0;
    int grid_div = 1;         // persistent grids are launched at 1/grid_div of their full size (the workers of a split batch share the SMs)
    struct QueueBufs { DevBuf node[BBQ_NODE_CLASSES][2], leaf[2], count; } qbuf[2];  // [0] normal, [1] wide-root reads


    cudaEvent_t ev[BB_N_STAGES + 1] = {};
    float stage_ms[BB_N_STAGES] = {};

    // sub-batches: a context made by bb_create owns n_kids further worker contexts on the same device; a batch is
    // dealt out over the workers (this context is worker 0) and their kernel chains run side by side on their
    // own streams, so that one worker's tails and host round trips are covered by the others' kernels
    std::vector<bb_ctx *> kids;
    int n_split = 1;                          // workers the uploaded batch is spread over (1: this context alone)
    std::vector<std::vector<int32_t>> part;   // part[w][i] = batch position of worker w's i-th read
    std::vector<int64_t> part_base;           // offset of worker w's block in the fetched seq / qual buffers
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;

    // launch trace (BADREAD_B200_TRACE=1): an event after every launch / host step of a run, dumped by bb_trace_dump
    bool trace = false;
    struct Mark { const char *name; int stream; cudaEvent_t ev; };
    std::vector<Mark> marks;
    std::vector<cudaEvent_t> mark_pool;
    size_t mark_used = 0;
};

// Records "the work enqueued on `st` up to here is done" under `name` (tracing only).
static void mark(bb_ctx *ctx, cudaStream_t st, const char *name) {
    if (!ctx->trace) return;
    if (ctx->mark_used == ctx->mark_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        ctx->mark_pool.push_back(e);
    }
    cudaEvent_t e = ctx->mark_pool[ctx->mark_used++];
    cudaEventRecord(e, st);
    int id = st == ctx->stream ? 0 : st == ctx->stream2 ? 1 : -1;
    for (int p = 0; p < 2 && id < 0; p++)
        for (int x = 0; x < 3; x++)
            if (st == ctx->side[p][x]) id = 2 + 4 * p + x;
    ctx->marks.push_back(bb_ctx::Mark{name, id < 0 ? 0 : id, e});
}

static thread_local std::string g_create_error;

// A context drives up to 7 streams per worker (4 workers by default).  CUDA multiplexes streams onto
// CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue serialize behind each other's
// pending waits.  Ask for the maximum unless the user chose a value; it only takes effect if CUDA is not initialized yet
// in this process (badread_b200/_lib.py and bench.py set it before anything touches CUDA).
namespace {
struct ConnectionsDefault {
    ConnectionsDefault() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_connections_default;
}  // namespace

#define BB_CUDA(ctx, call)                                                                                   \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
            return BB_ERR_CUDA;                                                                              \
        }                                                                                                    \
    } while (0)

static int set_err(bb_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return code;
}

extern "C" const char *bb_version(void) { return "badread_b200 0.1.0 (sm_100a)"; }

extern "C" const char *bb_last_error(const bb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" const char *bb_stage_name(int stage) {
    return (stage >= 0 && stage < BB_N_STAGES) ? kStageNames[stage] : "";
}

extern "C" int64_t bb_launch_count(const bb_ctx *ctx) {
    if (!ctx) return 0;
    int64_t total = ctx->launches;
    for (const bb_ctx *kid : ctx->kids) total += kid->launches;
    return total;
}

static int create_worker(bb_ctx **out, int device, uint64_t seed, bool high_priority = false) {
    if (!out) return BB_ERR_ARG;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev <= 0) {
        g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                         " (badread_b200 has no CPU path)";
        return BB_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) { g_create_error = "invalid device ordinal"; return BB_ERR_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return BB_ERR_CUDA; }
    bb_ctx *ctx = new bb_ctx();
    ctx->device = device;
    ctx->seed = seed;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return BB_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    // worker 0 carries the longest reads of a split batch (bb_batch_upload): their dependent chain of stages bounds the
    // step from below, so its kernels go first whenever the block scheduler has a choice
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const int prio = high_priority ? prio_hi : prio_lo;
    e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return BB_ERR_CUDA; }
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, prio);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_scan, cudaEventDisableTiming);
    for (int p = 0; p < 2; p++) {
        cudaEventCreateWithFlags(&ctx->ev_level[p], cudaEventDisableTiming);
        for (int x = 0; x < 3; x++) {
            cudaStreamCreateWithPriority(&ctx->side[p][x], cudaStreamNonBlocking, prio);
            cudaEventCreateWithFlags(&ctx->ev_side[p][x], cudaEventDisableTiming);
        }
    }
    if (cudaHostAlloc((void **)&ctx->h_info, sizeof(bb_ctx::RunInfo), cudaHostAllocPortable) != cudaSuccess) {
        g_create_error = "cudaHostAlloc failed"; delete ctx; return BB_ERR_CUDA;
    }
    // misc.REV_COMP_DICT (misc.py:56-61); anything else complements to 'N' (misc.py:64-68)
    uint8_t comp[256];
    std::memset(comp, 'N', sizeof(comp));
    const char *from = "ATGCatgcRYSWKMBVDHNryswkmbvdhn.-?";
    const char *to = "TACGtacgYRSWMKVBHDNyrswmkvbhdn.-?";
    for (int i = 0; from[i]; i++) comp[(uint8_t)from[i]] = (uint8_t)to[i];
    e = cudaMemcpyToSymbol(bb_c_comp, comp, 256);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return BB_ERR_CUDA; }
    e = bbl_node_pair_init();
    if (e == cudaSuccess) e = bbl_node_quad_init();
    if (e == cudaSuccess) e = bbl_window_lane_init();
    if (e == cudaSuccess) e = bbl_leaf_lane_init();
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return BB_ERR_CUDA; }
    // persistent warps: 4 CTAs of 4 warps per SM for the warp-per-read kernels
    ctx->n_warps = ctx->sm_count * 4 * BB_WARPS_PER_CTA;
    if (const char *e = std::getenv("BADREAD_B200_TRACE")) ctx->trace = (e[0] == '1');
    if (const char *e = std::getenv("BADREAD_B200_LANE8_COLS")) ctx->lane8_cols = std::atoi(e);
    if (const char *e = std::getenv("BADREAD_B200_PAIR_CTAS")) ctx->pair_ctas = (e[0] == '2') ? 2 : 1;
    if (const char *e = std::getenv("BADREAD_B200_HEAD_WORKER")) ctx->head_worker = (e[0] != '0');
    if (const char *e = std::getenv("BADREAD_B200_GRID_DIV")) ctx->grid_div_env = std::atoi(e);
    if (const char *e = std::getenv("BADREAD_B200_QUAD")) ctx->use_quad = (e[0] != '0');
    if (const char *e = std::getenv("BADREAD_B200_LOWMEM")) ctx->lowmem = (e[0] != '0');
    if (const char *e = std::getenv("BADREAD_B200_LPT")) ctx->lpt_order = (e[0] != '0');
    if (const char *e = std::getenv("BADREAD_B200_RING_T")) ctx->ring_t = (e[0] == '8') ? 8 : (e[0] == '2') ? 2 : 4;
    *out = ctx;
    return BB_OK;
}

extern "C" int bb_destroy(bb_ctx *ctx);
static void (*nccl_destroy)(void *) = nullptr;  // set once NCCL is loaded (bb_comm_init_*)

extern "C" int bb_create(bb_ctx **out, int device, uint64_t seed) {
    bool prio = true;
    if (const char *e = std::getenv("BADREAD_B200_HEAD_PRIORITY")) prio = (e[0] != '0');
    int rc = create_worker(out, device, seed, prio);
    if (rc) return rc;
    bb_ctx *ctx = *out;
    int n_workers = 4;
    if (const char *e = std::getenv("BADREAD_B200_SUBBATCHES")) n_workers = std::max(1, std::min(8, std::atoi(e)));
    for (int w = 1; w < n_workers; w++) {
        bb_ctx *kid = nullptr;
        if ((rc = create_worker(&kid, device, seed))) { bb_destroy(ctx); *out = nullptr; return rc; }
        ctx->kids.push_back(kid);
    }
    cudaEventCreate(&ctx->ev_t0);
    cudaEventCreate(&ctx->ev_t1);
    return BB_OK;
}

extern "C" int bb_destroy(bb_ctx *ctx) {
    if (!ctx) return BB_OK;
    for (bb_ctx *kid : ctx->kids) bb_destroy(kid);
    ctx->kids.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    DevBuf *bufs[] = {&ctx->ref, &ctx->em_k2r, &ctx->em_rowoff, &ctx->em_cum, &ctx->em_flags, &ctx->em_slots,
                      &ctx->em_pool, &ctx->em_rowinfo, &ctx->qm_hkeys, &ctx->qm_hvals, &ctx->qm_rowoff, &ctx->qm_scores, &ctx->qm_cum,
                      &ctx->d_read_index, &ctx->d_seg_off, &ctx->d_segs, &ctx->d_lit, &ctx->d_target, &ctx->d_order,
                      &ctx->d_reads, &ctx->d_kidx, &ctx->d_frag, &ctx->d_state, &ctx->d_seq, &ctx->d_ops, &ctx->d_dcnt,
                      &ctx->d_qual, &ctx->d_out_seq, &ctx->d_out_qual, &ctx->d_counter, &ctx->s_hist, &ctx->s_hbuf,
                      &ctx->s_lr, &ctx->s_stack, &ctx->s_tbuf, &ctx->s_peq, &ctx->s_ltbuf, &ctx->d_ctime, &ctx->d_chlog, &ctx->d_wres,
                      &ctx->d_wtasks, &ctx->d_wfallback, &ctx->d_active,
                      &ctx->d_fpeq, &ctx->d_speq, &ctx->d_fallback, &ctx->s_leafhist, &ctx->s_lr_lean, &ctx->s_wckpt, &ctx->s_lanehist, &ctx->d_scan, &ctx->d_red,
                      &ctx->p_q, &ctx->p_t, &ctx->p_ops, &ctx->p_dcnt, &ctx->p_out, &ctx->p_qual};
    for (auto &qb : ctx->qbuf) {
        for (auto &cl : qb.node) for (auto &d : cl) d.release();
        qb.leaf[0].release(); qb.leaf[1].release(); qb.count.release();
    }
    for (DevBuf *b : bufs) b->release();
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_scan) cudaEventDestroy(ctx->ev_scan);
    for (int p = 0; p < 2; p++) {
        if (ctx->ev_level[p]) cudaEventDestroy(ctx->ev_level[p]);
        for (int x = 0; x < 3; x++) {
            if (ctx->side[p][x]) cudaStreamDestroy(ctx->side[p][x]);
            if (ctx->ev_side[p][x]) cudaEventDestroy(ctx->ev_side[p][x]);
        }
    }
    if (ctx->h_info) cudaFreeHost(ctx->h_info);
    if (ctx->nccl_comm && nccl_destroy) nccl_destroy(ctx->nccl_comm);
    for (cudaEvent_t e : ctx->mark_pool) cudaEventDestroy(e);
    delete ctx;
    return BB_OK;
}

template <typename T>
static int upload(bb_ctx *ctx, DevBuf &buf, const T *src, size_t count) {
    BB_CUDA(ctx, buf.ensure(std::max<size_t>(count, 1) * sizeof(T)));
    if (count) BB_CUDA(ctx, cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return BB_OK;
}

extern "C" int bb_upload_reference(bb_ctx *ctx, const uint8_t *bases, int64_t n_bases) {
    if (!ctx || n_bases < 0 || (n_bases && !bases)) return set_err(ctx, BB_ERR_ARG, "bb_upload_reference: bad arguments");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = upload(ctx, ctx->ref, bases, (size_t)n_bases);
    if (rc) return rc;
    BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->ref_len = n_bases;
    for (bb_ctx *kid : ctx->kids) { kid->ref.borrow(ctx->ref); kid->ref_len = n_bases; }  // one copy per GPU
    return BB_OK;
}

extern "C" int bb_upload_error_model(bb_ctx *ctx, int k, int type, const int32_t *kmer_to_row, int64_t n_index,
                                     int32_t n_rows, const int32_t *row_off, const double *cum, const uint8_t *flags,
                                     const uint32_t *slots, const uint8_t *pool, int64_t pool_len) {
    if (!ctx) return BB_ERR_ARG;
    if (k < 1 || k > 12 || (type != 0 && type != 1)) return set_err(ctx, BB_ERR_ARG, "error model: k must be 1..12");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->em = BBErrorModelDev{};
    ctx->em.k = k; ctx->em.type = type;
    if (type == 1) {
        if (!kmer_to_row || !row_off || !cum || !flags || !slots || n_rows <= 0 || n_index != (1ll << (2 * k)))
            return set_err(ctx, BB_ERR_ARG, "error model: missing tables");
        const int64_t ne = row_off[n_rows];
        int rc;
        if ((rc = upload(ctx, ctx->em_k2r, kmer_to_row, (size_t)n_index))) return rc;
        if ((rc = upload(ctx, ctx->em_rowoff, row_off, (size_t)n_rows + 1))) return rc;
        if ((rc = upload(ctx, ctx->em_cum, cum, (size_t)ne))) return rc;
        if ((rc = upload(ctx, ctx->em_flags, flags, (size_t)ne))) return rc;
        if ((rc = upload(ctx, ctx->em_slots, slots, (size_t)ne * k))) return rc;
        if ((rc = upload(ctx, ctx->em_pool, pool, (size_t)pool_len))) return rc;
        std::vector<BBRowInfo> info((size_t)n_rows);
        for (int32_t r = 0; r < n_rows; r++) {
            const int32_t e0 = row_off[r], ne = row_off[r + 1] - e0;
            if (ne <= 0) return set_err(ctx, BB_ERR_ARG, "error model: empty table row");
            BBRowInfo &ri = info[(size_t)r];
            ri.cum_last = cum[e0 + ne - 1]; ri.cum0 = cum[e0]; ri.e0 = e0; ri.ne = ne;
            ri.first_is_identity = flags[e0] == 1 ? 1 : 0; ri.pad = 0;
        }
        if ((rc = upload(ctx, ctx->em_rowinfo, info.data(), info.size()))) return rc;
        BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `info` is about to go out of scope
        ctx->em.kmer_to_row = ctx->em_k2r.as<int32_t>(); ctx->em.row_off = ctx->em_rowoff.as<int32_t>();
        ctx->em.cum = ctx->em_cum.as<double>(); ctx->em.flags = ctx->em_flags.as<uint8_t>();
        ctx->em.slots = ctx->em_slots.as<uint32_t>(); ctx->em.pool = ctx->em_pool.as<uint8_t>();
        ctx->em.rowinfo = ctx->em_rowinfo.as<BBRowInfo>();
    }
    BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_em = true;
    ctx->uploaded = false;
    for (bb_ctx *kid : ctx->kids) { kid->em = ctx->em; kid->have_em = true; kid->uploaded = false; }  // shared tables
    return BB_OK;
}

extern "C" int bb_upload_qscore_model(bb_ctx *ctx, int kmer_size, int32_t n_keys, const uint64_t *keys,
                                      const int32_t *row_off, const uint8_t *scores, const double *cum) {
    if (!ctx) return BB_ERR_ARG;
    if (kmer_size < 1 || (kmer_size & 1) == 0 || n_keys <= 0 || !keys || !row_off || !scores || !cum)
        return set_err(ctx, BB_ERR_ARG, "qscore model: bad arguments");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t bits = 6;
    while ((1ull << bits) < 2ull * (uint64_t)n_keys) bits++;
    const size_t hsize = (size_t)1 << bits;
    std::vector<uint64_t> hk(hsize, 0);
    std::vector<int32_t> hv(hsize, -1);
    for (int32_t i = 0; i < n_keys; i++) {
        if (keys[i] < 4) return set_err(ctx, BB_ERR_ARG, "qscore model: invalid packed key");
        uint32_t h = (uint32_t)((keys[i] * 0x9E3779B97F4A7C15ull) >> (64 - bits));
        while (hk[h] != 0 && hk[h] != keys[i]) h = (h + 1) & (uint32_t)(hsize - 1);
        hk[h] = keys[i]; hv[h] = i;  // a repeated key keeps its last row, like the dict assignment in load_from_file
    }
    const int64_t ne = row_off[n_keys];
    int rc;
    if ((rc = upload(ctx, ctx->qm_hkeys, hk.data(), hsize))) return rc;
    if ((rc = upload(ctx, ctx->qm_hvals, hv.data(), hsize))) return rc;
    if ((rc = upload(ctx, ctx->qm_rowoff, row_off, (size_t)n_keys + 1))) return rc;
    if ((rc = upload(ctx, ctx->qm_scores, scores, (size_t)ne))) return rc;
    if ((rc = upload(ctx, ctx->qm_cum, cum, (size_t)ne))) return rc;
    BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->qm.kmer_size = kmer_size; ctx->qm.hbits = bits;
    ctx->qm.hkeys = ctx->qm_hkeys.as<uint64_t>(); ctx->qm.hvals = ctx->qm_hvals.as<int32_t>();
    ctx->qm.row_off = ctx->qm_rowoff.as<int32_t>(); ctx->qm.scores = ctx->qm_scores.as<uint8_t>();
    ctx->qm.cum = ctx->qm_cum.as<double>();
    ctx->have_qm = true;
    for (bb_ctx *kid : ctx->kids) { kid->qm = ctx->qm; kid->have_qm = true; }  // shared tables
    return BB_OK;
}

// Scratch shared by the warp-per-read kernels. hist is sized for the largest traceback edlib's 1 MiB rule
// admits (ceil(n/64)*m < 52429 -> < 104858 32-row blocks); tbuf holds a joined 1000-slot window.
static int ensure_scratch(bb_ctx *ctx, int hbuf_need, int lr_need, int len_need) {
    const int n_warps = ctx->n_warps;
    const int hist_cap = 106496;
    const int tbuf_stride = 1000 * 255 + 1024;
    const int stack_cap = 64;
    int hbuf_cap = std::max(hbuf_need + 64, tbuf_stride);
    hbuf_cap = (hbuf_cap + 255) & ~255;
    int lr_cap = std::max(lr_need + 64, 4096);
    lr_cap = (lr_cap + 255) & ~255;
    int peq_cap = bb_peq_words(len_need) + 8;
    peq_cap = (peq_cap + 63) & ~63;
    if (ctx->pool.peq_cap >= peq_cap) peq_cap = ctx->pool.peq_cap;
    if (ctx->pool.hbuf_cap >= hbuf_cap) hbuf_cap = ctx->pool.hbuf_cap;
    if (ctx->pool.lr_cap >= lr_cap) lr_cap = ctx->pool.lr_cap;
    BB_CUDA(ctx, ctx->s_hist.ensure((size_t)n_warps * hist_cap * sizeof(uint2)));
    BB_CUDA(ctx, ctx->s_tbuf.ensure((size_t)n_warps * tbuf_stride));
    BB_CUDA(ctx, ctx->s_stack.ensure((size_t)n_warps * stack_cap * 5 * sizeof(int)));
    BB_CUDA(ctx, ctx->s_hbuf.ensure((size_t)n_warps * hbuf_cap));
    BB_CUDA(ctx, ctx->s_lr.ensure((size_t)n_warps * lr_cap * 2 * sizeof(int)));
    BB_CUDA(ctx, ctx->s_peq.ensure((size_t)n_warps * peq_cap * sizeof(uint4)));
    BBScratchPool &p = ctx->pool;
    p.hist = ctx->s_hist.as<uint2>(); p.hist_stride = hist_cap; p.hist_cap = hist_cap;
    p.hbuf = ctx->s_hbuf.as<int8_t>(); p.hbuf_stride = hbuf_cap; p.hbuf_cap = hbuf_cap;
    p.lr = ctx->s_lr.as<int>(); p.lr_stride = 2ll * lr_cap; p.lr_cap = lr_cap;
    p.stack = ctx->s_stack.as<int>(); p.stack_cap = stack_cap;
    p.tbuf = ctx->s_tbuf.as<uint8_t>(); p.tbuf_stride = tbuf_stride;
    p.peq = ctx->s_peq.as<uint4>(); p.peq_stride = peq_cap; p.peq_cap = peq_cap;
    // the single-warp node kernels (2 pipelines x 3 widths, all resident at once) only touch the split-score arrays
    constexpr int lean_cap = 2048;  // > a + b + 1 of the widest lean class (bb_pick_L<4>(a, b, 16) > 0: a + b < 1920)
    const size_t lean_warps = 2 * (size_t)ctx->sm_count * (4 + 6 + 6) * BB_WARPS_PER_CTA;  // room for the grid knobs' maxima
    BB_CUDA(ctx, ctx->s_lr_lean.ensure(lean_warps * lean_cap * 2 * sizeof(int)));
    ctx->pool_lean = p;
    ctx->pool_lean.lr = ctx->s_lr_lean.as<int>(); ctx->pool_lean.lr_stride = 2ll * lean_cap; ctx->pool_lean.lr_cap = lean_cap;
    return BB_OK;
}

static BBBatchDev batch_dev(bb_ctx *ctx) {
    BBBatchDev B{};
    B.n_reads = ctx->n_reads;
    B.read_index = ctx->d_read_index.as<unsigned long long>();
    B.seg_off = ctx->d_seg_off.as<int>();
    B.segs = ctx->d_segs.as<bb_segment>();
    B.lit = ctx->d_lit.as<uint8_t>();
    B.target = ctx->d_target.as<double>();
    B.order = ctx->d_order.as<int>();
    B.reads = ctx->d_reads.as<BBReadDev>();
    B.frag = ctx->d_frag.as<uint8_t>();
    B.state = ctx->d_state.as<uint32_t>();
    B.kidx = ctx->d_kidx.as<int>();
    B.seq = ctx->d_seq.as<uint8_t>();
    B.ops = ctx->d_ops.as<uint8_t>();
    B.dcnt = ctx->d_dcnt.as<unsigned int>();
    B.qual = ctx->d_qual.as<uint8_t>();
    B.out_seq = ctx->d_out_seq.as<uint8_t>();
    B.out_qual = ctx->d_out_qual.as<uint8_t>();
    B.fpeq = ctx->d_fpeq.as<uint4>();
    B.speq = ctx->d_speq.as<uint4>();
    B.ctime = ctx->d_ctime.as<unsigned int>();
    B.chlog = ctx->d_chlog.as<uint2>();
    B.wres = ctx->d_wres.as<int2>();
    return B;
}

// Counters of a run live in d_counter (256 ints, cleared once per run): [0, 16) spare; round r of the error loop owns
// the 16 ints from BB_ROUND_BASE(r).
#define BB_N_COUNTERS 256
#define BB_MAX_ROUNDS 15
#define BB_ROUND_BASE(r) (16 + 16 * (r))
enum { BBC_MUTATE = 0, BBC_NTASKS = 1, BBC_LANE4 = 2, BBC_FB1 = 3, BBC_LANE8 = 4, BBC_FB2 = 5, BBC_WARP = 6, BBC_PENDING = 7 };

// Hirschberg levels a fragment of max_len bases can need: a node is split while edlib's traceback estimate
// 20 ceil(nn/64) mm + 8 mm reaches 1 MiB; mm halves per level and nn <= 2 mm + 64 bounds the query side generously.
static int level_bound(int max_len, double slack) {
    long long mm = (long long)(max_len * slack) + 64;
    int d = 0;
    while (20ll * ((2 * mm + 64 + 63) / 64) * mm + 8ll * mm >= 1048576ll) { mm = (mm + 1) / 2; d++; }
    return d + 1;
}

// Every allocation a run needs, sized from the fragment lengths: a run is pure enqueueing, nothing on the host
// depends on a value the device computes.  What turns out too small is flagged by the kernels and w_finish grows
// the knobs (slack, n_rounds, levels, lr_worst) and runs the batch again.
static int w_prepare(bb_ctx *ctx) {
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = ctx->n_reads;
    const int64_t off = ctx->frag_total;
    ctx->seq_cap = (int64_t)((double)off * ctx->slack) + 16ll * n + 65536;
    ctx->out_cap = ctx->seq_cap;
    ctx->speq_cap = ctx->seq_cap / 32 + (2ll * BB_PEQ_PAD + 2) * n + 64;
    BB_CUDA(ctx, ctx->d_frag.ensure((size_t)off + 16));
    BB_CUDA(ctx, ctx->d_state.ensure(((size_t)off + 16) * sizeof(uint32_t)));
    BB_CUDA(ctx, ctx->d_kidx.ensure(((size_t)off + 16) * sizeof(int)));
    BB_CUDA(ctx, ctx->d_counter.ensure(BB_N_COUNTERS * sizeof(int)));
    BB_CUDA(ctx, ctx->d_scan.ensure(sizeof(BBScanOut)));
    BB_CUDA(ctx, ctx->d_fpeq.ensure(((size_t)ctx->fpeq_total + 4) * sizeof(uint4)));
    BB_CUDA(ctx, ctx->d_ctime.ensure(((size_t)off + 16) * sizeof(unsigned int)));
    BB_CUDA(ctx, ctx->d_chlog.ensure(((size_t)ctx->log_total + 16) * sizeof(uint2)));
    BB_CUDA(ctx, ctx->d_wres.ensure(((size_t)ctx->wres_total + 16) * sizeof(int2)));
    BB_CUDA(ctx, ctx->d_wtasks.ensure(((size_t)ctx->wres_total + 16) * sizeof(BBWinTask)));
    BB_CUDA(ctx, ctx->d_wfallback.ensure((2 * (size_t)ctx->wres_total + 16) * sizeof(BBWinTask)));
    BB_CUDA(ctx, ctx->d_seq.ensure((size_t)ctx->seq_cap + 16));
    BB_CUDA(ctx, ctx->d_ops.ensure((size_t)ctx->seq_cap + 16));
    BB_CUDA(ctx, ctx->d_dcnt.ensure(((size_t)ctx->seq_cap + 16) * sizeof(unsigned int)));
    BB_CUDA(ctx, ctx->d_qual.ensure((size_t)ctx->seq_cap + 16));
    BB_CUDA(ctx, ctx->d_speq.ensure(((size_t)ctx->speq_cap + 4) * sizeof(uint4)));
    BB_CUDA(ctx, ctx->d_out_seq.ensure((size_t)ctx->out_cap + 16));
    BB_CUDA(ctx, ctx->d_out_qual.ensure((size_t)ctx->out_cap + 16));
    const int lane_ctas = ctx->sm_count * 4;  // 64-thread CTAs of the lane kernels
    {   // lane pools of the window aligner and the leaf aligner (the two alignment pipelines each own half)
        const size_t lanes = (size_t)lane_ctas * 64;
        if (ctx->lowmem) BB_CUDA(ctx, ctx->s_leafhist.ensure(2 * lanes * BB_LEAF_MAX_TILES * BB_LEAF_CKPT_WORDS * sizeof(uint32_t)));
        else BB_CUDA(ctx, ctx->s_lanehist.ensure(2 * lanes * BB_LEAF_LANE_COLS * BB_LEAF_LW * sizeof(uint2)));  // per-column history
        BB_CUDA(ctx, ctx->s_ltbuf.ensure(2 * lanes * BB_WIN_MAX_COLS));  // the 4-word window kernel runs up to 1.5x the lanes
        // window aligners: a checkpoint (2 LW + 2 words) per 16 columns per lane instead of a per-column history
        if (ctx->lowmem) BB_CUDA(ctx, ctx->s_wckpt.ensure(2 * lanes * BB_WIN_MAX_TILES * BB_WIN_CKPT_WORDS(BB_WIN_LW) * sizeof(uint32_t)));
    }
    // per-warp scratch: strip carries / bitmaps for the longest joined read; split-score arrays for the widest band
    // (expected: a few times the injected edits; worst case: the whole read)
    const int len_b = (int)std::min<double>((double)ctx->max_len * ctx->slack + 64.0, (double)(1 << 24));
    int lr_need = 4096;
    for (int r = 0; r < n; r++) {
        const double len = (double)ctx->h_reads[(size_t)r].frag_len;
        const double worst = len * ctx->slack + 64.0;
        const double expect = 3.0 * (1.0 - ctx->h_target[(size_t)r]) * len + 0.02 * len + 512.0;
        lr_need = std::max(lr_need, (int)std::min(worst, ctx->lr_worst ? worst : expect));
    }
    int rc = ensure_scratch(ctx, len_b, lr_need, len_b);
    if (rc) return rc;
    const int cap_node = (int)std::min<int64_t>(ctx->seq_cap / 256 + 4ll * n + 1024, 0x7ffffff0);
    for (int s = 0; s < 2; s++) {
        auto &qb = ctx->qbuf[s];
        for (int c = 0; c < BBQ_NODE_CLASSES; c++)
            for (int p = 0; p < 2; p++) BB_CUDA(ctx, qb.node[c][p].ensure((size_t)cap_node * sizeof(BBNode)));
        for (int w = 0; w < 2; w++) BB_CUDA(ctx, qb.leaf[w].ensure((size_t)cap_node * sizeof(BBNode)));
        BB_CUDA(ctx, qb.count.ensure(512 * sizeof(int)));
    }
    ctx->n_levels = std::min(48, level_bound(ctx->max_len, ctx->slack) + ctx->extra_levels);
    return BB_OK;
}

static int w_batch_upload(bb_ctx *ctx, int32_t n_reads, const uint64_t *read_index, const int32_t *seg_off,
                          const bb_segment *segs, const uint8_t *literal_pool, int64_t literal_len,
                          const double *target_identity) {
    if (!ctx) return BB_ERR_ARG;
    if (n_reads <= 0 || !read_index || !seg_off || !segs || !target_identity || literal_len < 0)
        return set_err(ctx, BB_ERR_ARG, "bb_batch_upload: bad arguments");
    if (!ctx->have_em || !ctx->have_qm) return set_err(ctx, BB_ERR_STATE, "upload the error and qscore models first");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int k = ctx->em.k;
    ctx->h_reads.assign((size_t)n_reads, BBReadDev{});
    ctx->h_inlen.assign((size_t)n_reads, 0);
    ctx->h_target.assign(target_identity, target_identity + n_reads);
    int64_t off = 0, peq_off = 0, log_off = 0, wres_off = 0;
    int max_len = 0;
    for (int32_t r = 0; r < n_reads; r++) {
        int64_t len = 0;
        if (seg_off[r + 1] < seg_off[r]) return set_err(ctx, BB_ERR_ARG, "seg_off must be non-decreasing");
        for (int32_t s = seg_off[r]; s < seg_off[r + 1]; s++) {
            const bb_segment &sg = segs[s];
            if (sg.len < 0 || sg.src < 0) return set_err(ctx, BB_ERR_ARG, "negative segment");
            if (sg.kind == BB_SEG_LITERAL) { if (sg.src + sg.len > literal_len) return set_err(ctx, BB_ERR_ARG, "literal segment out of range"); }
            else if (sg.kind == BB_SEG_REF_FWD || sg.kind == BB_SEG_REF_REV) { if (sg.src + sg.len > ctx->ref_len) return set_err(ctx, BB_ERR_ARG, "reference segment out of range"); }
            else return set_err(ctx, BB_ERR_ARG, "unknown segment kind");
            len += sg.len;
        }
        if (len + 2 * k >= (1 << 24)) return set_err(ctx, BB_ERR_ARG, "fragment too long (16 Mb limit)");
        ctx->h_inlen[(size_t)r] = (int32_t)len;
        BBReadDev &rd = ctx->h_reads[(size_t)r];
        rd.frag_off = off;
        rd.frag_len = (int)(len + 2 * k);
        rd.fpeq_off = peq_off;
        peq_off += bb_peq_words(rd.frag_len);
        {   // speculative loop bookkeeping: the loop cannot apply more than 0.9*len + k changes (simulate.py:285)
            const int cap = (int)(0.9 * (double)rd.frag_len) + k + 2;
            rd.log_off = log_off; rd.wres_off = wres_off;
            log_off += cap; wres_off += cap / BB_ALIGNMENT_INTERVAL + 1;
            const double need = (double)rd.frag_len * (1.0 - target_identity[r]);
            rd.horizon = (int)std::min<double>((double)cap, std::max(0.0, 1.25 * need) + 48.0);
            rd.n_logged = 0; rd.n_resume = 0; rd.a_done = 0; rd.status = BB_READ_PENDING; rd.stop_reason = 0;
        }
        off += (rd.frag_len + 15) & ~15;
        max_len = std::max(max_len, rd.frag_len);
    }
    ctx->frag_total = off; ctx->fpeq_total = peq_off; ctx->log_total = log_off; ctx->wres_total = wres_off;
    ctx->max_len = max_len;
    std::vector<int> order((size_t)n_reads);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int x, int y) { return ctx->h_reads[(size_t)x].frag_len > ctx->h_reads[(size_t)y].frag_len; });
    ctx->n_reads = n_reads;
    ctx->h_order = order;
    int rc;
    if ((rc = upload(ctx, ctx->d_read_index, read_index, (size_t)n_reads))) return rc;
    if ((rc = upload(ctx, ctx->d_seg_off, seg_off, (size_t)n_reads + 1))) return rc;
    if ((rc = upload(ctx, ctx->d_segs, segs, (size_t)seg_off[n_reads]))) return rc;
    if ((rc = upload(ctx, ctx->d_lit, literal_pool, (size_t)literal_len))) return rc;
    if ((rc = upload(ctx, ctx->d_target, target_identity, (size_t)n_reads))) return rc;
    if ((rc = upload(ctx, ctx->d_order, order.data(), (size_t)n_reads))) return rc;
    if ((rc = upload(ctx, ctx->d_reads, ctx->h_reads.data(), (size_t)n_reads))) return rc;
    if ((rc = w_prepare(ctx))) return rc;
    BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->uploaded = true;
    ctx->ran = false;
    ctx->finished = false;
    return BB_OK;
}

// Persistent grids (CTAs that pull work from a queue until it is empty) of `per_sm` CTAs per SM at full size.  The
// workers of a split batch run side by side: each launches its share, so that their kernels are resident together
// instead of queueing behind each other's long-running CTAs.
static int pgrid(const bb_ctx *ctx, int per_sm, const char *knob = nullptr) {
    if (knob) {  // tuning: BADREAD_B200_GRID_<KNOB> = CTAs per SM of that kernel
        static thread_local char name[64];
        std::snprintf(name, sizeof(name), "BADREAD_B200_GRID_%s", knob);
        if (const char *e = std::getenv(name)) { const int v = std::atoi(e); if (v > 0) per_sm = v; }
    }
    return std::max(ctx->sm_count / 2, (ctx->sm_count * per_sm + ctx->grid_div - 1) / ctx->grid_div);
}

// The error loop decoupled from its identity re-measurements (bb_loop.cuh): mutate ahead -> task list -> all window
// alignments as independent lane tasks -> scalar replay, n_rounds times back to back.  A round after the last read
// has finished costs six launches that find nothing to do; a read that is still pending after the last round is
// reported by the replay kernel's counter and w_finish runs the batch again with more rounds.
static int enqueue_error_loop(bb_ctx *ctx, const BBBatchDev &B) {
    cudaStream_t st = ctx->stream;
    const int n = ctx->n_reads;
    int *cnt = ctx->d_counter.as<int>();
    const int lane_ctas = ctx->sm_count * 4;
    const int *order = ctx->d_order.as<int>();
    BBWinTask *tasks = ctx->d_wtasks.as<BBWinTask>();
    BBWinTask *fb1 = ctx->d_wfallback.as<BBWinTask>(), *fb2 = fb1 + ctx->wres_total + 8;
    for (int round = 0; round < ctx->n_rounds; round++) {
        int *c = cnt + BB_ROUND_BASE(round);
        bbl_mutate(std::min(pgrid(ctx, 8, "MUTATE"), n), st, B, ctx->em, ctx->seed, c + BBC_MUTATE, order, n, ctx->is_head);
        mark(ctx, st, "mutate");
        bb_k_window_tasks<<<(n + 255) / 256, 256, 0, st>>>(B, order, n, tasks, c + BBC_NTASKS);
        mark(ctx, st, "window_tasks");
        // 4-word windows first (bands up to 64 rows: almost every window); what does not fit falls through to the
        // 8-word build and from there to the warp kernel
        if (ctx->lowmem)
            bbl_window_lane4(std::min(pgrid(ctx, 6, "WIN4"), ctx->sm_count * 8), st, B, ctx->em, tasks, c + BBC_NTASKS, ctx->seed,
                             ctx->s_wckpt.as<uint32_t>(), ctx->s_ltbuf.as<uint8_t>(), c + BBC_LANE4, fb1, c + BBC_FB1);
        else
            bbl_window_lane_hist(4, ctx->ring_t, std::min(pgrid(ctx, 8, "WIN4"), ctx->sm_count * 8), st, B, ctx->em, tasks, c + BBC_NTASKS,
                                 ctx->seed, ctx->s_lanehist.as<uint2>(), ctx->s_ltbuf.as<uint8_t>(), c + BBC_LANE4, fb1, c + BBC_FB1);
        mark(ctx, st, "window_lane4");
        if (ctx->lowmem)
            bbl_window_lane8(std::min(pgrid(ctx, 3, "WIN8"), ctx->sm_count * 8), st, B, ctx->em, fb1, c + BBC_FB1, ctx->seed,
                             ctx->s_wckpt.as<uint32_t>(), ctx->s_ltbuf.as<uint8_t>(), c + BBC_LANE8, fb2, c + BBC_FB2);
        else
            bbl_window_lane_hist(8, 4, std::min(pgrid(ctx, 4, "WIN8"), ctx->sm_count * 4), st, B, ctx->em, fb1, c + BBC_FB1, ctx->seed,
                                 ctx->s_lanehist.as<uint2>(), ctx->s_ltbuf.as<uint8_t>(), c + BBC_LANE8, fb2, c + BBC_FB2);
        mark(ctx, st, "window_lane8");
        bbl_window_warp(pgrid(ctx, 2), st, B, ctx->em, ctx->pool, fb2, c + BBC_FB2, ctx->seed, c + BBC_WARP);
        mark(ctx, st, "window_warp");
        bb_k_replay<<<(n + 3) / 4, 128, 0, st>>>(B, order, n, ctx->em.k, c + BBC_PENDING);
        mark(ctx, st, "replay");
        ctx->launches += 6;
    }
    return BB_OK;
}

// Final alignment as level-synchronous tasks (bb_tasks.cuh): every level of all reads' Hirschberg trees is a few
// launches (warp-pair, lean-warp and lane nodes), leaves run at the end.  The number of levels comes from the longest
// fragment; nodes left over after the last level are reported by the queue counters (w_finish adds levels).
static int enqueue_align_tasks(bb_ctx *ctx, const BBBatchDev &B) {
    cudaStream_t stream[2] = {ctx->stream, ctx->stream2};
    const int n = ctx->n_reads;
    const int cap_node = (int)std::min<int64_t>(ctx->seq_cap / 256 + 4ll * n + 1024, 0x7ffffff0);
    const int lane_ctas = ctx->sm_count * 4;
    const size_t hist_per_pipe = (size_t)lane_ctas * 64 * BB_LEAF_MAX_TILES * BB_LEAF_CKPT_WORDS;  // checkpoint words
    BBQueues Q[2];
    int *cnt[2];
    for (int s = 0; s < 2; s++) {
        auto &qb = ctx->qbuf[s];
        for (int c = 0; c < BBQ_NODE_CLASSES; c++)
            for (int p = 0; p < 2; p++) Q[s].node[c][p] = qb.node[c][p].as<BBNode>();
        for (int w = 0; w < 2; w++) Q[s].leaf[w] = qb.leaf[w].as<BBNode>();
        cnt[s] = qb.count.as<int>();
        Q[s].count = cnt[s]; Q[s].overflow = cnt[s] + BBQ_OVERFLOW; Q[s].cap_node = cap_node; Q[s].cap_leaf = cap_node;
        Q[s].lane8_cols = ctx->lane8_cols;
        BB_CUDA(ctx, cudaMemsetAsync(cnt[s], 0, 512 * sizeof(int), stream[0]));
    }
    bb_k_push_roots<<<(n + 255) / 256, 256, 0, stream[0]>>>(B, Q[0], Q[1], ctx->d_order.as<int>());
    ctx->launches++;
    // pipeline 0 (stream 0): every read whose root band fits the lean / lane kernels; pipeline 1 (stream 1): reads
    // with a wide root (long or noisy reads).  The two never wait for each other's levels.
    BB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, stream[0]));
    BB_CUDA(ctx, cudaStreamWaitEvent(stream[1], ctx->ev_fork, 0));
    int *cursor[2] = {cnt[0] + 16, cnt[1] + 16};
    const int warp_base[2] = {0, ctx->n_warps / 2};
    const int w4 = ctx->sm_count * 4 * BB_WARPS_PER_CTA, w2 = ctx->sm_count * 6 * BB_WARPS_PER_CTA;  // scratch slots of the lean kernels (maxima)
    const int lean_base[2] = {0, w4 + 2 * w2};
    // The node classes of a level read the same queues and push into the next level's: they are independent and run
    // side by side on their own streams; the level ends when all of them have finished.
    for (int level = 0; level < ctx->n_levels; level++) {
        // (the roots are queued longest read first; every later queue fills in the order the parents finish, longest
        // last, and is walked from its end)
        const int p = (level & 1) | (level > 0 && ctx->lpt_order ? BBQ_BACKWARDS : 0);
        for (int s = 0; s < 2; s++) {
            cudaStream_t st = stream[s];
            BB_CUDA(ctx, cudaMemsetAsync(cnt[s] + BBQ_COUNT(0, (p & 1) ^ 1), 0, BBQ_NODE_CLASSES * sizeof(int), st));
            BB_CUDA(ctx, cudaEventRecord(ctx->ev_level[s], st));
            int n_side = 0;
            auto on_side = [&]() -> cudaStream_t {
                cudaStream_t x = ctx->side[s][n_side++];
                cudaStreamWaitEvent(x, ctx->ev_level[s], 0);
                mark(ctx, x, "fork");
                return x;
            };
            if (s == 1) {
                cudaStream_t x = on_side();
                if (ctx->use_quad) bbl_node_quad(ctx->sm_count, x, B, Q[s], ctx->pool, p, cursor[s]++, warp_base[s]);
                else bbl_node_pair(ctx->sm_count * ctx->pair_ctas, x, B, Q[s], ctx->pool, p, cursor[s]++, warp_base[s]);
                ctx->launches++;
                mark(ctx, x, ctx->use_quad ? "node_quad" : "node_pair");
            }
            {
                cudaStream_t x = on_side();   // the two narrow single-warp classes share a stream
                bbl_node_warp(2, std::min(pgrid(ctx, 3, "WARP2"), ctx->sm_count * 6), x, B, Q[s], ctx->pool_lean, p, cursor[s]++, lean_base[s] + w4);
                mark(ctx, x, "node_warp2");
                bbl_node_warp(1, std::min(pgrid(ctx, 3, "WARP1"), ctx->sm_count * 6), x, B, Q[s], ctx->pool_lean, p, cursor[s]++, lean_base[s] + w4 + w2);
                mark(ctx, x, "node_warp1");
                x = on_side();
                bbl_node_lane8(pgrid(ctx, 6, "LANE8"), x, B, Q[s], p, cursor[s]++);
                mark(ctx, x, "node_lane8");
            }
            bbl_node_warp(4, std::min(pgrid(ctx, 2, "WARP4"), ctx->sm_count * 4), st, B, Q[s], ctx->pool_lean, p, cursor[s]++, lean_base[s]);
            mark(ctx, st, "node_warp4");
            ctx->launches += 4;
            for (int x = 0; x < n_side; x++) {
                BB_CUDA(ctx, cudaEventRecord(ctx->ev_side[s][x], ctx->side[s][x]));
                BB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_side[s][x], 0));
            }
        }
    }
    for (int s = 0; s < 2; s++) {
        cudaStream_t st = stream[s];
        bbl_leaf_warp(ctx->sm_count, st, B, Q[s], ctx->pool, cursor[s]++, warp_base[s]);
        mark(ctx, st, "leaf_warp");
        if (ctx->lowmem)
            bbl_leaf_lane(std::min(pgrid(ctx, 3, "LEAF"), ctx->sm_count * 4), st, B, Q[s], ctx->s_leafhist.as<uint32_t>() + s * hist_per_pipe, cursor[s]++);
        else
            bbl_leaf_lane_hist(std::min(pgrid(ctx, 4, "LEAF"), ctx->sm_count * 4), st, B, Q[s],
                               ctx->s_lanehist.as<uint2>() + s * ((size_t)lane_ctas * 64 * BB_LEAF_LANE_COLS * BB_LEAF_LW), cursor[s]++);
        mark(ctx, st, "leaf_lane");
        ctx->launches += 2;
    }
    BB_CUDA(ctx, cudaEventRecord(ctx->ev_join, stream[1]));
    BB_CUDA(ctx, cudaStreamWaitEvent(stream[0], ctx->ev_join, 0));
    for (int s = 0; s < 2; s++)
        BB_CUDA(ctx, cudaMemcpyAsync(ctx->h_info->qcount[s], cnt[s], 32 * sizeof(int), cudaMemcpyDeviceToHost, stream[0]));
    return BB_OK;
}

// Enqueues the whole hot path of the uploaded batch on the worker's streams and returns: no host round trip inside.
static int w_enqueue(bb_ctx *ctx) {
    if (!ctx) return BB_ERR_ARG;
    if (!ctx->uploaded) return set_err(ctx, BB_ERR_STATE, "bb_batch_run: no batch uploaded");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = ctx->n_reads;
    BBBatchDev B = batch_dev(ctx);
    ctx->finished = false;
    // the per-read records start from the uploaded state on every run (bb_batch_run may be repeated)
    BB_CUDA(ctx, cudaMemcpyAsync(ctx->d_reads.p, ctx->h_reads.data(), (size_t)n * sizeof(BBReadDev), cudaMemcpyHostToDevice, st));
    BB_CUDA(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, BB_N_COUNTERS * sizeof(int), st));
    ctx->marks.clear(); ctx->mark_used = 0;
    mark(ctx, st, "begin");
    BB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    bb_k_build_fragments<<<n, 256, 0, st>>>(B, ctx->ref.as<uint8_t>(), ctx->em.k, ctx->seed,
                                             ctx->em.type == 1 ? ctx->em.kmer_to_row : nullptr);
    ctx->launches++;
    mark(ctx, st, "build_fragments");
    BB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    int rc = enqueue_error_loop(ctx, B);
    if (rc) return rc;
    BB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    // offsets of the per-read regions of the joined reads, on the device
    bb_k_scan<<<1, 1024, 0, st>>>(B, n, ctx->seq_cap, ctx->out_cap, ctx->speq_cap, ctx->d_scan.as<BBScanOut>());
    ctx->launches++;
    BB_CUDA(ctx, cudaMemcpyAsync(&ctx->h_info->scan, ctx->d_scan.p, sizeof(BBScanOut), cudaMemcpyDeviceToHost, st));
    BB_CUDA(ctx, cudaEventRecord(ctx->ev_scan, st));  // from here on the host can learn the size of this worker's output
    BB_CUDA(ctx, cudaMemsetAsync(ctx->d_dcnt.p, 0, ((size_t)ctx->seq_cap + 16) * sizeof(unsigned int), st));
    mark(ctx, st, "scan");
    BB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    bb_k_join<<<n, 256, 0, st>>>(B, ctx->em);
    ctx->launches++;
    mark(ctx, st, "join");
    BB_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
    if ((rc = enqueue_align_tasks(ctx, B))) return rc;
    mark(ctx, st, "align_tail");
    BB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    bb_k_qscores<<<n, 256, 0, st>>>(B, ctx->qm, ctx->seed);
    ctx->launches++;
    mark(ctx, st, "qscores");
    BB_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
    bb_k_compact<<<n, 256, 0, st>>>(B);
    ctx->launches++;
    mark(ctx, st, "compact");
    BB_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
    BB_CUDA(ctx, cudaMemcpyAsync(ctx->h_info->counters, ctx->d_counter.p, BB_N_COUNTERS * sizeof(int), cudaMemcpyDeviceToHost, st));
    BB_CUDA(ctx, cudaGetLastError());
    ctx->ran = true;
    return BB_OK;
}

// Waits for the worker's run and checks what the device reported.  Returns BB_OK when the results are final; when
// something did not fit (buffers sized from the fragment lengths, rounds, levels, split-score scratch) the knob is
// raised and the batch runs again - the results do not depend on any of them.
static int w_finish(bb_ctx *ctx) {
    if (!ctx) return BB_ERR_ARG;
    if (!ctx->ran) return set_err(ctx, BB_ERR_STATE, "no run to finish");
    if (ctx->finished) return BB_OK;
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = ctx->n_reads;
    for (int attempt = 0;; attempt++) {
        ctx->h_res.resize((size_t)n);
        BB_CUDA(ctx, cudaMemcpyAsync(ctx->h_res.data(), ctx->d_reads.p, (size_t)n * sizeof(BBReadDev), cudaMemcpyDeviceToHost, ctx->stream));
        BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const bb_ctx::RunInfo &info = *ctx->h_info;
        bool again = false;
        std::string why;
        if (info.counters[BB_ROUND_BASE(ctx->n_rounds - 1) + BBC_PENDING] > 0 || info.scan.n_pending > 0) {
            ctx->n_rounds = std::min(BB_MAX_ROUNDS, ctx->n_rounds + 3); again = true; why += " error-loop rounds";
        }
        if (info.scan.n_nospace > 0) {  // (reads still pending after the last round are counted separately)
            const double need = (double)std::max(info.scan.seq_total, info.scan.out_total) / (double)std::max<int64_t>(1, ctx->frag_total);
            ctx->slack = std::max(ctx->slack * 1.5, need * 1.1 + 0.05); again = true; why += " buffer slack";
        }
        const int last_parity = ctx->n_levels & 1;  // the queues the level after the last one would read
        int left = 0, overflow = 0;
        for (int s = 0; s < 2; s++) {
            for (int c = 0; c < BBQ_NODE_CLASSES; c++) left += info.qcount[s][BBQ_COUNT(c, last_parity)];
            overflow += info.qcount[s][BBQ_OVERFLOW];
        }
        if (left > 0) { ctx->extra_levels += 8; again = true; why += " levels"; }
        if (overflow) { ctx->slack *= 1.5; again = true; why += " task queues"; }
        for (int r = 0; r < n && !ctx->lr_worst; r++) {
            const int f = (ctx->h_res[(size_t)r].flags & ~BB_FLAG_NOSPACE) >> 8;
            if (f & (16 | 4 | 2)) { ctx->lr_worst = true; ctx->slack *= 1.25; again = true; why += " alignment scratch"; }
        }
        if (!again) break;
        ctx->reran = true;
        if (attempt >= 3) return set_err(ctx, BB_ERR_INTERNAL, "batch did not fit after growing:" + why);
        int rc = w_prepare(ctx);
        if (rc) return rc;
        if ((rc = w_enqueue(ctx))) return rc;
    }
    ctx->out_total = ctx->h_info->scan.out_total;
    ctx->finished = true;
    return BB_OK;
}

extern "C" int bb_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return BB_ERR_ARG;
    *ptr = nullptr;
    return cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable) == cudaSuccess ? BB_OK : BB_ERR_CUDA;
}

extern "C" int bb_host_free(void *ptr) {
    if (!ptr) return BB_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? BB_OK : BB_ERR_CUDA;
}

extern "C" int bb_synchronize(bb_ctx *ctx) {
    if (!ctx) return BB_ERR_ARG;
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (bb_ctx *kid : ctx->kids) BB_CUDA(ctx, cudaStreamSynchronize(kid->stream));
    return BB_OK;
}

static int w_last_run_ms(bb_ctx *ctx, float *total_ms, float *stage_ms) {
    if (!ctx) return BB_ERR_ARG;
    if (!ctx->ran) return set_err(ctx, BB_ERR_STATE, "no run to time");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    BB_CUDA(ctx, cudaEventSynchronize(ctx->ev[BB_N_STAGES - 1]));
    for (int i = 0; i < BB_N_STAGES - 1; i++)
        BB_CUDA(ctx, cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]));
    BB_CUDA(ctx, cudaEventElapsedTime(&ctx->stage_ms[BB_N_STAGES - 1], ctx->ev[0], ctx->ev[BB_N_STAGES - 1]));
    if (total_ms) *total_ms = ctx->stage_ms[BB_N_STAGES - 1];
    if (stage_ms) std::memcpy(stage_ms, ctx->stage_ms, sizeof(ctx->stage_ms));
    return BB_OK;
}

// Enqueues the device-to-host copies of a finished (or at least scanned) worker's packed block on its stream.
static int w_copy_out(bb_ctx *ctx, int64_t base, uint8_t *seq_out, uint8_t *qual_out) {
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t total = ctx->h_info->scan.out_total;
    if (total > 0) {
        if (!seq_out || !qual_out) return set_err(ctx, BB_ERR_ARG, "null output buffers");
        BB_CUDA(ctx, cudaMemcpyAsync(seq_out + base, ctx->d_out_seq.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
        BB_CUDA(ctx, cudaMemcpyAsync(qual_out + base, ctx->d_out_qual.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return BB_OK;
}

// results[pos[i]] describes the worker's i-th read, its out_off shifted by `base` (from the records w_finish fetched).
static int w_results(bb_ctx *ctx, bb_read_result *results, const int32_t *pos, int64_t base) {
    const int n = ctx->n_reads;
    int bad = 0, bad_read = -1;
    for (int r = 0; r < n; r++) {
        const BBReadDev &rd = ctx->h_res[(size_t)r];
        if (results) {
            bb_read_result &o = results[pos ? pos[r] : r];
            o.out_off = base + rd.out_off; o.out_len = rd.out_len; o.frag_len = ctx->h_inlen[(size_t)r];
            o.matches = rd.matches; o.columns = rd.seq_len + rd.dels; o.loop_count = rd.loop_count;
            o.change_count = rd.change_count; o.n_alignments = rd.n_align; o.flags = rd.flags;
            o.loop_kcycles = rd.kc_loop; o.align_kcycles = rd.kc_align;
        }
        if (rd.flags && !bad) { bad = rd.flags; bad_read = r; }
    }
    if (bad) {
        char msg[160];
        std::snprintf(msg, sizeof(msg), "device invariant violated: read %d flags 0x%x", pos ? pos[bad_read] : bad_read, bad);
        return set_err(ctx, BB_ERR_INTERNAL, msg);
    }
    return BB_OK;
}

// ---- batch entry points: deal the reads out over the workers ---------------------------------------------
static bb_ctx *worker_of(bb_ctx *ctx, int w) { return w == 0 ? ctx : ctx->kids[(size_t)w - 1]; }

extern "C" int bb_batch_upload(bb_ctx *ctx, int32_t n_reads, const uint64_t *read_index, const int32_t *seg_off,
                               const bb_segment *segs, const uint8_t *literal_pool, int64_t literal_len,
                               const double *target_identity) {
    if (!ctx) return BB_ERR_ARG;
    if (n_reads <= 0 || !read_index || !seg_off || !segs || !target_identity || literal_len < 0)
        return set_err(ctx, BB_ERR_ARG, "bb_batch_upload: bad arguments");
    const int n_workers = 1 + (int)ctx->kids.size();
    ctx->n_split = (n_workers > 1 && n_reads >= 64 * n_workers) ? n_workers : 1;
    if (ctx->n_split == 1) {
        ctx->is_head = false;
        ctx->grid_div = 1;
        return w_batch_upload(ctx, n_reads, read_index, seg_off, segs, literal_pool, literal_len, target_identity);
    }
    // deal the reads out longest first, so that every worker sees the same length distribution
    const int S = ctx->n_split;
    std::vector<int64_t> len((size_t)n_reads, 0);
    for (int32_t r = 0; r < n_reads; r++) {
        if (seg_off[r + 1] < seg_off[r]) return set_err(ctx, BB_ERR_ARG, "seg_off must be non-decreasing");
        for (int32_t x = seg_off[r]; x < seg_off[r + 1]; x++) len[(size_t)r] += segs[x].len;
    }
    std::vector<int32_t> order((size_t)n_reads);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return len[(size_t)x] > len[(size_t)y]; });
    ctx->part.assign((size_t)S, std::vector<int32_t>());
    // The chain of Hirschberg levels of the longest reads is the longest dependent chain of the batch.  Worker 0 is a
    // small HEAD batch of the longest reads: its error loop is short, so their alignment starts early and overlaps the
    // other workers' error loops instead of trailing the step.  The other workers share the rest evenly.
    int32_t n_head = 0;
    if (S >= 3 && ctx->head_worker) {
        int64_t total = 0, hb = 0;
        for (int32_t r = 0; r < n_reads; r++) total += len[(size_t)r];
        const int64_t longest = len[(size_t)order[0]];
        while (n_head < n_reads / 16 && hb < total / S &&
               (10 * len[(size_t)order[(size_t)n_head]] >= 6 * longest || hb < total / (4 * S))) {
            hb += len[(size_t)order[(size_t)n_head]];
            ctx->part[0].push_back(order[(size_t)n_head++]);
        }
        if (n_head < 16) { ctx->part[0].clear(); n_head = 0; }
    }
    ctx->is_head = n_head > 0;
    {   // share of the SMs each worker's persistent kernels ask for (BADREAD_B200_GRID_DIV overrides; 1 = all of them)
        int div = ctx->grid_div_env > 0 ? ctx->grid_div_env : 1;
        for (int w = 0; w < S; w++) worker_of(ctx, w)->grid_div = div;
    }
    if (n_head > 0) for (int32_t i = n_head; i < n_reads; i++) ctx->part[(size_t)(1 + (i - n_head) % (S - 1))].push_back(order[(size_t)i]);
    else for (int32_t i = 0; i < n_reads; i++) ctx->part[(size_t)(i % S)].push_back(order[(size_t)i]);
    constexpr int kBadLiteral = 1 << 20;  // not a bb_status value
    std::vector<int> rcs((size_t)S, 0);
    std::vector<std::thread> threads;
    auto upload_part = [&](int w) {
        std::vector<int32_t> &mine = ctx->part[(size_t)w];
        std::sort(mine.begin(), mine.end());
        std::vector<uint64_t> ridx; std::vector<int32_t> soff(1, 0); std::vector<bb_segment> sg; std::vector<uint8_t> lit;
        std::vector<double> ident;
        for (int32_t r : mine) {
            ridx.push_back(read_index[r]);
            ident.push_back(target_identity[r]);
            for (int32_t x = seg_off[r]; x < seg_off[r + 1]; x++) {
                bb_segment g = segs[x];
                if (g.kind == BB_SEG_LITERAL) {
                    if (g.len < 0 || g.src < 0 || g.src + g.len > literal_len) { rcs[(size_t)w] = kBadLiteral; return; }
                    const int64_t at = (int64_t)lit.size();
                    lit.insert(lit.end(), literal_pool + g.src, literal_pool + g.src + g.len);
                    g.src = at;
                }
                sg.push_back(g);
            }
            soff.push_back((int32_t)sg.size());
        }
        rcs[(size_t)w] = w_batch_upload(worker_of(ctx, w), (int32_t)mine.size(), ridx.data(), soff.data(), sg.data(),
                                        lit.data(), (int64_t)lit.size(), ident.data());
    };
    for (int w = 1; w < S; w++) threads.emplace_back(upload_part, w);
    upload_part(0);
    for (auto &t : threads) t.join();
    for (int w = 0; w < S; w++) {
        if (rcs[(size_t)w] == kBadLiteral) return set_err(ctx, BB_ERR_ARG, "literal segment out of range");
        if (rcs[(size_t)w]) return w == 0 ? rcs[0] : set_err(ctx, rcs[(size_t)w], worker_of(ctx, w)->err);
    }
    return BB_OK;
}

// Asynchronous: the kernel chains of all workers are enqueued from this thread and overlap on the device.
extern "C" int bb_batch_run(bb_ctx *ctx) {
    if (!ctx) return BB_ERR_ARG;
    const int S = ctx->n_split;
    for (int w = 0; w < S; w++) worker_of(ctx, w)->reran = false;
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    BB_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    for (int w = 1; w < S; w++) BB_CUDA(ctx, cudaStreamWaitEvent(worker_of(ctx, w)->stream, ctx->ev_t0, 0));
    for (int w = 0; w < S; w++) {
        const int rc = w_enqueue(worker_of(ctx, w));
        if (rc) return w == 0 ? rc : set_err(ctx, rc, worker_of(ctx, w)->err);
    }
    for (int w = 1; w < S; w++) BB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, worker_of(ctx, w)->ev[BB_N_STAGES - 1], 0));
    BB_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
    return BB_OK;
}

// Whole-batch device time from the first worker's first kernel to the last worker's last one.  With several
// workers the stages of different workers overlap in time: each stage is reported as its share of the summed
// per-worker stage times, scaled to the whole-batch time.
extern "C" int bb_last_run_ms(bb_ctx *ctx, float *total_ms, float *stage_ms) {
    if (!ctx) return BB_ERR_ARG;
    const int S = ctx->n_split;
    float sum[BB_N_STAGES] = {};
    for (int w = 0; w < S; w++) {
        float t = 0.f, st[BB_N_STAGES];
        const int rc = w_last_run_ms(worker_of(ctx, w), &t, st);
        if (rc) return w == 0 ? rc : set_err(ctx, rc, worker_of(ctx, w)->err);
        for (int i = 0; i < BB_N_STAGES; i++) sum[i] += st[i];
    }
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    BB_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
    float total = 0.f;
    BB_CUDA(ctx, cudaEventElapsedTime(&total, ctx->ev_t0, ctx->ev_t1));
    const float scale = sum[BB_N_STAGES - 1] > 0.f ? total / sum[BB_N_STAGES - 1] : 0.f;
    for (int i = 0; i < BB_N_STAGES - 1; i++) ctx->stage_ms[i] = sum[i] * scale;
    ctx->stage_ms[BB_N_STAGES - 1] = total;
    if (total_ms) *total_ms = total;
    if (stage_ms) std::memcpy(stage_ms, ctx->stage_ms, sizeof(ctx->stage_ms));
    return BB_OK;
}

// Launch trace of the last run (BADREAD_B200_TRACE=1) as CSV: worker, stream, name, begin_ms, end_ms relative to the
// first worker's first mark; "begin" is the previous mark on the same worker and stream.
extern "C" int bb_trace_dump(bb_ctx *ctx, const char *path) {
    if (!ctx || !path) return BB_ERR_ARG;
    if (!ctx->trace || ctx->marks.empty()) return set_err(ctx, BB_ERR_STATE, "no trace (set BADREAD_B200_TRACE=1 before bb_create)");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    BB_CUDA(ctx, cudaDeviceSynchronize());
    FILE *f = std::fopen(path, "w");
    if (!f) return set_err(ctx, BB_ERR_ARG, "cannot open trace file");
    std::fprintf(f, "worker,stream,name,begin_ms,end_ms\n");
    const cudaEvent_t base = ctx->marks[0].ev;
    const int S = ctx->n_split;
    for (int w = 0; w < S; w++) {
        bb_ctx *wk = w == 0 ? ctx : ctx->kids[(size_t)w - 1];
        float prev[10] = {};
        bool have[10] = {};
        for (const bb_ctx::Mark &m : wk->marks) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, base, m.ev) != cudaSuccess) continue;
            const float b = have[m.stream] ? prev[m.stream] : (have[0] ? prev[0] : t);
            std::fprintf(f, "%d,%d,%s,%.4f,%.4f\n", w, m.stream, m.name, b, t);
            prev[m.stream] = t; have[m.stream] = true;
        }
    }
    std::fclose(f);
    return BB_OK;
}

// Finishes every worker's run, packs their blocks back to back in the caller's buffers and fills the results.
// eager: the copies of a worker were already enqueued behind its kernels with these bases (bb_sequence_batch).
static int fetch_all(bb_ctx *ctx, bb_read_result *results, uint8_t *seq_out, uint8_t *qual_out, int64_t out_cap,
                     int64_t *out_total, const std::vector<int64_t> *eager_bases) {
    const int S = ctx->n_split;
    std::vector<int64_t> base((size_t)S, 0);
    int64_t total = 0;
    bool moved = false;  // a worker's output size changed after the eager copies were placed (it had to run again)
    for (int w = 0; w < S; w++) {
        bb_ctx *wk = worker_of(ctx, w);
        if (!wk->ran) return set_err(ctx, BB_ERR_STATE, "bb_fetch_last_batch: nothing to fetch");
        const int rc = w_finish(wk);
        if (rc) return w == 0 ? rc : set_err(ctx, rc, wk->err);
        if (wk->reran) moved = true;
        base[(size_t)w] = total;
        total += wk->out_total;
    }
    ctx->part_base = base;
    if (out_total) *out_total = total;
    if (out_cap < total) return set_err(ctx, BB_ERR_CAPACITY, "output buffers too small");
    if (total && (!seq_out || !qual_out)) return set_err(ctx, BB_ERR_ARG, "null output buffers");
    const bool have_eager = eager_bases && !moved && *eager_bases == base;
    if (!have_eager) {
        for (int w = 0; w < S; w++) {
            const int rc = w_copy_out(worker_of(ctx, w), base[(size_t)w], seq_out, qual_out);
            if (rc) return w == 0 ? rc : set_err(ctx, rc, worker_of(ctx, w)->err);
        }
    }
    for (int w = 0; w < S; w++) {
        bb_ctx *wk = worker_of(ctx, w);
        BB_CUDA(ctx, cudaSetDevice(wk->device));
        BB_CUDA(ctx, cudaStreamSynchronize(wk->stream));
        const int rc = w_results(wk, results, S == 1 ? nullptr : ctx->part[(size_t)w].data(), base[(size_t)w]);
        if (rc) return w == 0 ? rc : set_err(ctx, rc, wk->err);
    }
    return BB_OK;
}

extern "C" int bb_fetch_last_batch(bb_ctx *ctx, bb_read_result *results, uint8_t *seq_out, uint8_t *qual_out,
                                   int64_t out_cap, int64_t *out_total) {
    if (!ctx) return BB_ERR_ARG;
    return fetch_all(ctx, results, seq_out, qual_out, out_cap, out_total, nullptr);
}

extern "C" int bb_sequence_batch(bb_ctx *ctx, int32_t n_reads, const uint64_t *read_index, const int32_t *seg_off,
                                 const bb_segment *segs, const uint8_t *literal_pool, int64_t literal_len,
                                 const double *target_identity, bb_read_result *results, uint8_t *seq_out,
                                 uint8_t *qual_out, int64_t out_cap, int64_t *out_total) {
    int rc = bb_batch_upload(ctx, n_reads, read_index, seg_off, segs, literal_pool, literal_len, target_identity);
    if (rc) return rc;
    if ((rc = bb_batch_run(ctx))) return rc;
    // each worker's block is copied out as soon as its own chain is done, while the others still compute: its place
    // in the caller's buffers only needs the output sizes of the workers before it, known since their scans
    const int S = ctx->n_split;
    std::vector<int64_t> base((size_t)S, 0);
    int64_t total = 0;
    bool eager = true;
    for (int w = 0; w < S && eager; w++) {
        bb_ctx *wk = worker_of(ctx, w);
        BB_CUDA(ctx, cudaSetDevice(wk->device));
        BB_CUDA(ctx, cudaEventSynchronize(wk->ev_scan));
        base[(size_t)w] = total;
        total += wk->h_info->scan.out_total;
        if (wk->h_info->scan.n_nospace > 0 || total > out_cap) { eager = false; break; }
        if ((rc = w_copy_out(wk, base[(size_t)w], seq_out, qual_out))) { eager = false; break; }
    }
    return fetch_all(ctx, results, seq_out, qual_out, out_cap, out_total, eager ? &base : nullptr);
}

// ---- single-pair entry points ------------------------------------------------------------------------
static int align_pair_device(bb_ctx *ctx, const uint8_t *q, int n, const uint8_t *t, int m, DevBuf &dq, DevBuf &dt,
                             DevBuf &dops, DevBuf &ddcnt, DevBuf &dout, int out5[5]) {
    int rc;
    if ((rc = upload(ctx, dq, q, (size_t)n))) return rc;
    if ((rc = upload(ctx, dt, t, (size_t)m))) return rc;
    BB_CUDA(ctx, dops.ensure((size_t)n + 16));
    BB_CUDA(ctx, ddcnt.ensure(((size_t)n + 16) * sizeof(unsigned int)));
    BB_CUDA(ctx, dout.ensure(8 * sizeof(int)));
    if ((rc = ensure_scratch(ctx, std::max(n, m), std::max(n, m), std::max(n, m)))) return rc;
    BB_CUDA(ctx, cudaMemsetAsync(ddcnt.p, 0, ((size_t)n + 16) * sizeof(unsigned int), ctx->stream));
    BB_CUDA(ctx, cudaMemsetAsync(dout.p, 0, 8 * sizeof(int), ctx->stream));
    bbl_align_pair(ctx->stream, dq.as<uint8_t>(), n, dt.as<uint8_t>(), m, std::max(n, m), ctx->pool, dops.as<uint8_t>(),
                   ddcnt.as<unsigned int>(), dout.as<int>());
    ctx->launches++;
    BB_CUDA(ctx, cudaMemcpyAsync(out5, dout.p, 5 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out5[4]) {
        char msg[96];
        std::snprintf(msg, sizeof(msg), "aligner invariant violated (code 0x%x)", out5[4]);
        return set_err(ctx, BB_ERR_INTERNAL, msg);
    }
    return BB_OK;
}

extern "C" int bb_align_path(bb_ctx *ctx, const uint8_t *query, int32_t q_len, const uint8_t *target, int32_t t_len,
                             uint8_t *ops_out, int64_t ops_cap, int64_t *n_ops, int32_t *distance) {
    if (!ctx) return BB_ERR_ARG;
    if (!query || !target || q_len <= 0 || t_len <= 0) return set_err(ctx, BB_ERR_ARG, "bb_align_path: empty sequence");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf &dq = ctx->p_q, &dt = ctx->p_t, &dops = ctx->p_ops, &ddcnt = ctx->p_dcnt, &dout = ctx->p_out;
    int out5[5] = {0, 0, 0, 0, 0};
    int rc = align_pair_device(ctx, query, q_len, target, t_len, dq, dt, dops, ddcnt, dout, out5);
    std::vector<uint8_t> ops((size_t)q_len);
    std::vector<unsigned int> dcnt((size_t)q_len);
    if (rc == BB_OK) {
        cudaMemcpy(ops.data(), dops.p, (size_t)q_len, cudaMemcpyDeviceToHost);
        cudaMemcpy(dcnt.data(), ddcnt.p, (size_t)q_len * sizeof(unsigned int), cudaMemcpyDeviceToHost);
    }
    if (rc) return rc;
    const int64_t total = (int64_t)q_len + out5[1];
    if (n_ops) *n_ops = total;
    if (distance) *distance = out5[2];
    if (total > ops_cap) return set_err(ctx, BB_ERR_CAPACITY, "ops buffer too small");
    static const char sym[3] = {'=', 'X', 'I'};
    int64_t w = 0;
    for (int x = 0; x < out5[3]; x++) ops_out[w++] = 'D';
    for (int i = 0; i < q_len; i++) {
        ops_out[w++] = (uint8_t)sym[ops[(size_t)i] < 3 ? ops[(size_t)i] : 0];
        for (unsigned int x = 0; x < dcnt[(size_t)i]; x++) ops_out[w++] = 'D';
    }
    if (w != total) return set_err(ctx, BB_ERR_INTERNAL, "column count mismatch");
    return BB_OK;
}

extern "C" int bb_get_qscores(bb_ctx *ctx, uint64_t read_index, const uint8_t *seq, int32_t seq_len,
                              const uint8_t *frag, int32_t frag_len, uint8_t *qual_out, int32_t *matches,
                              int32_t *columns) {
    if (!ctx) return BB_ERR_ARG;
    if (!seq || !frag || seq_len <= 0 || frag_len <= 0 || !qual_out) return set_err(ctx, BB_ERR_ARG, "bb_get_qscores: bad arguments");
    if (!ctx->have_qm) return set_err(ctx, BB_ERR_STATE, "upload the qscore model first");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf &dq = ctx->p_q, &dt = ctx->p_t, &dops = ctx->p_ops, &ddcnt = ctx->p_dcnt, &dout = ctx->p_out, &dqual = ctx->p_qual;
    int out5[5] = {0, 0, 0, 0, 0};
    int rc = align_pair_device(ctx, seq, seq_len, frag, frag_len, dq, dt, dops, ddcnt, dout, out5);
    if (rc == BB_OK) {
        cudaError_t e = dqual.ensure((size_t)seq_len + 16);
        if (e != cudaSuccess) rc = set_err(ctx, BB_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc == BB_OK) {
        bb_k_qscores_pair<<<(seq_len + 255) / 256, 256, 0, ctx->stream>>>(dops.as<uint8_t>(), ddcnt.as<unsigned int>(), seq_len,
                                                                            ctx->qm, ctx->seed, read_index, dqual.as<uint8_t>());
        ctx->launches++;
        cudaMemcpyAsync(qual_out, dqual.p, (size_t)seq_len, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = set_err(ctx, BB_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc) return rc;
    if (matches) *matches = out5[0];
    if (columns) *columns = seq_len + out5[1];
    return BB_OK;
}

// ---- the one collective of the path: SUM of emitted bases over the GPUs (stop condition, simulate.py:63) ----------
// Reads shard over GPUs by read index and never exchange data; the only thing the GPUs have to agree on is the running
// total of emitted bases that ends the simulation.  NCCL is loaded at run time (the process's own libnccl.so.2 if one
// is already mapped - e.g. PyTorch's - else the system library), so the library has no link-time dependency on it.
namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, bb_nccl_id, int) = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // NCCL writes its banner / warnings to stdout unless told otherwise: stdout is where the FASTQ goes
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return api;
    api.lib = h;
    api.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void **, int, bb_nccl_id, int))dlsym(h, "ncclCommInitRank");
    api.CommInitAll = (int (*)(void **, int, const int *))dlsym(h, "ncclCommInitAll");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclAllReduce");
    api.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    api.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommAbort");  // teardown must not wait for peers that already left
    if (!api.CommDestroy) api.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    nccl_destroy = (void (*)(void *))api.CommDestroy;
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.AllReduce && api.GroupStart && api.GroupEnd;
    return api;
}
constexpr int kNcclInt64 = 4, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)

int nccl_err(bb_ctx *ctx, const char *what, int rc) {
    const char *msg = nccl().GetErrorString ? nccl().GetErrorString(rc) : "?";
    return set_err(ctx, BB_ERR_CUDA, std::string(what) + ": " + msg);
}
}  // namespace

extern "C" int bb_nccl_available(void) { return nccl().ok ? 1 : 0; }

// NCCL prints its version banner with a plain printf to stdout when NCCL_DEBUG=VERSION (init.cc showVersion; the debug
// file setting does not apply to it) - and stdout is where `badread simulate` writes the FASTQ (measured: 29 bytes that
// made the 2-GPU output differ from the 1-GPU one).  While a communicator is created, file descriptor 1 points at stderr.
struct StdoutToStderr {
    int saved = -1;
    StdoutToStderr() {
        fflush(stdout);
        saved = dup(1);
        if (saved >= 0) dup2(2, 1);
    }
    ~StdoutToStderr() {
        fflush(stdout);
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};

extern "C" int bb_comm_unique_id(bb_nccl_id *id) {
    if (!id) return BB_ERR_ARG;
    if (!nccl().ok) return BB_ERR_STATE;
    return nccl().GetUniqueId(id) == 0 ? BB_OK : BB_ERR_CUDA;
}

extern "C" int bb_comm_init_rank(bb_ctx *ctx, const bb_nccl_id *id, int rank, int world) {
    if (!ctx || !id || rank < 0 || rank >= world) return BB_ERR_ARG;
    if (!nccl().ok) return set_err(ctx, BB_ERR_STATE, "libnccl.so.2 could not be loaded");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->nccl_comm && nccl().CommDestroy) { nccl().CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    int rc;
    {
        StdoutToStderr guard;
        rc = nccl().CommInitRank(&ctx->nccl_comm, world, *id, rank);
    }
    if (rc) return nccl_err(ctx, "ncclCommInitRank", rc);
    BB_CUDA(ctx, ctx->d_red.ensure(2 * sizeof(long long)));
    return BB_OK;
}

extern "C" int bb_comm_init_all(bb_ctx **ctxs, int n) {
    if (!ctxs || n <= 0) return BB_ERR_ARG;
    if (!nccl().ok) return set_err(ctxs[0], BB_ERR_STATE, "libnccl.so.2 could not be loaded");
    std::vector<int> devs((size_t)n);
    std::vector<void *> comms((size_t)n, nullptr);
    for (int i = 0; i < n; i++) { if (!ctxs[i]) return BB_ERR_ARG; devs[(size_t)i] = ctxs[i]->device; }
    int rc;
    {
        StdoutToStderr guard;
        rc = nccl().CommInitAll(comms.data(), n, devs.data());
    }
    if (rc) return nccl_err(ctxs[0], "ncclCommInitAll", rc);
    for (int i = 0; i < n; i++) {
        ctxs[i]->nccl_comm = comms[(size_t)i];
        BB_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        BB_CUDA(ctxs[i], ctxs[i]->d_red.ensure(2 * sizeof(long long)));
    }
    return BB_OK;
}

// One process per GPU: every rank passes its local count, all get the sum.
extern "C" int bb_allreduce_bases(bb_ctx *ctx, int64_t local, int64_t *total) {
    if (!ctx || !total) return BB_ERR_ARG;
    if (!ctx->nccl_comm) return set_err(ctx, BB_ERR_STATE, "bb_allreduce_bases: no communicator (bb_comm_init_rank)");
    BB_CUDA(ctx, cudaSetDevice(ctx->device));
    long long *d = ctx->d_red.as<long long>();
    const long long v = local;
    BB_CUDA(ctx, cudaMemcpyAsync(d, &v, sizeof(v), cudaMemcpyHostToDevice, ctx->stream));
    const int rc = nccl().AllReduce(d, d + 1, 1, kNcclInt64, kNcclSum, ctx->nccl_comm, ctx->stream);
    if (rc) return nccl_err(ctx, "ncclAllReduce", rc);
    long long out = 0;
    BB_CUDA(ctx, cudaMemcpyAsync(&out, d + 1, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
    BB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *total = out;
    return BB_OK;
}

// One process, several GPUs (the CLI's --gpus N): the contexts' counts are summed in one NCCL group call.
extern "C" int bb_allreduce_bases_all(bb_ctx **ctxs, int n, const int64_t *local, int64_t *total) {
    if (!ctxs || n <= 0 || !local || !total) return BB_ERR_ARG;
    for (int i = 0; i < n; i++)
        if (!ctxs[i] || !ctxs[i]->nccl_comm) return set_err(ctxs[0], BB_ERR_STATE, "bb_allreduce_bases_all: no communicator (bb_comm_init_all)");
    for (int i = 0; i < n; i++) {
        BB_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        const long long v = local[i];
        BB_CUDA(ctxs[i], cudaMemcpyAsync(ctxs[i]->d_red.p, &v, sizeof(v), cudaMemcpyHostToDevice, ctxs[i]->stream));
        BB_CUDA(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));  // `v` leaves scope
    }
    int rc = nccl().GroupStart();
    for (int i = 0; i < n && !rc; i++) {
        long long *d = ctxs[i]->d_red.as<long long>();
        rc = nccl().AllReduce(d, d + 1, 1, kNcclInt64, kNcclSum, ctxs[i]->nccl_comm, ctxs[i]->stream);
    }
    const int rc2 = nccl().GroupEnd();
    if (rc || rc2) return nccl_err(ctxs[0], "ncclAllReduce (group)", rc ? rc : rc2);
    long long out = 0;
    BB_CUDA(ctxs[0], cudaSetDevice(ctxs[0]->device));
    BB_CUDA(ctxs[0], cudaMemcpyAsync(&out, ctxs[0]->d_red.as<long long>() + 1, sizeof(out), cudaMemcpyDeviceToHost, ctxs[0]->stream));
    for (int i = 0; i < n; i++) {
        BB_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        BB_CUDA(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
    }
    *total = out;
    return BB_OK;
}
