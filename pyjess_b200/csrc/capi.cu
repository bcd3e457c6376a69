// capi.cu -- host side of the C ABI declared in include/jess_b200.h.
//
// Owns the HBM layout (DESIGN.md "Data layout"), de-duplicates template-atom
// specs, drives the kernels of kernels.cuh chunk by chunk on ONE stream with a
// grow-and-retry protocol for the frontier/candidate/hit buffers, and returns
// canonically ordered hits.  No CPU fallback: without a CUDA device every entry
// point fails with JB_ERR_CUDA.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/jess_b200.h"
#include "kernels.cuh"
#include "expand_lv.cuh"
#include "expand_grid.cuh"
#include "expand_dfs.cuh"
#include "hits.cuh"
#include "order.cuh"
#include "prepare_small.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges are no-ops unless a profiler is attached

// NVTX range around one stage of a run (tracing): visible to nsys / ncu --nvtx.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

using namespace jb;

// ---- errors -------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(e_ == cudaErrorMemoryAllocation ? JB_ERR_NOMEM : JB_ERR_CUDA, "%s: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                             \
    } while (0)

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- device buffers ---------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    size_t recent = 0;        // largest request since the last context_decay
    // grow-only; contents are NOT preserved unless keep > 0 (bytes to copy over)
    cudaError_t ensure(size_t need, cudaStream_t st, size_t keep = 0)
    {
        recent = std::max(recent, need);
        if (need <= bytes) return cudaSuccess;
        size_t nb = std::max(need, bytes + bytes / 2);
        nb = (nb + 255) & ~(size_t)255;
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, nb);
        if (e != cudaSuccess) return e;
        if (keep && p) {
            e = cudaMemcpyAsync(q, p, std::min(keep, bytes), cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) { cudaFree(q); return e; }
            cudaStreamSynchronize(st);
        }
        if (p) cudaFree(p);
        p = q;
        bytes = nb;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; recent = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {   // grow-only page-locked host staging
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        size_t nb = std::max(need, bytes + bytes / 2);
        void *q = nullptr;
        cudaError_t e = cudaMallocHost(&q, nb);
        if (e != cudaSuccess) return e;
        if (p) cudaFreeHost(p);
        p = q;
        bytes = nb;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

template <class T> static cudaError_t upload(DevBuf &b, const T *src, size_t n, cudaStream_t st)
{
    cudaError_t e = b.ensure(std::max<size_t>(n, 1) * sizeof(T), st);
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    return cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
}

// ---- handles ---------------------------------------------------------------------
struct jb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    size_t ws_limit = 0;          // bytes the frontier buffers may use in total
    int group = 2;                // lanes per partial match (JB_GROUP)
    int group_lv[kMaxN + 1] = {};  // per-level override (JB_GROUPS="g1,g2,..." for levels 1, 2, ...); 0: default
    int debug = 0;                // JB_DEBUG: per-level counts and times on stderr
    int legacy = 0;               // JB_LEGACY: run-time-level k_expand at every level (A/B comparisons)
    int sweep = 1;                // JB_SWEEP: 1 cell-grid sweep (expand_grid.cuh), 0 x-range stream (expand_lv.cuh)
    long long dfs_pairs = 4096;   // JB_DFS_PAIRS: chunks of at most that many pairs take the one-launch depth-first sweep
    int first_slice_pct = 20;     // JB_FIRST_SLICE: see query_host_pipelined (0: equal slices)
    int grid_levels = 24;         // JB_GRID_LEVELS: levels 1 .. this take the cell-grid kernel (kMaxGridLv), deeper ones k_expand<G>
    long long tail_items = 0;     // JB_TAIL_ITEMS: a frontier of at most that many partial matches is finished depth first (k_expand_tail); 0 = off
    cudaEvent_t lev[kMaxN + 1] = {};
    // workspace (grow only)
    DevBuf crec, chead, cxtab, cbin, cbinlist, recs, ordbuf, fr[2], tuples, ctr;
    DevBuf h_mol, h_tpl, h_atoms, h_rmsd, h_rot, h_c0, h_c1;   // hits as k_superpose appends them
    DevBuf f_mol, f_tpl, f_atoms, f_rmsd, f_rot, f_c0, f_c1;   // ... in canonical order (hits.cuh)
    DevBuf f_xyz;                                              // JB_QUERY_COORDS: transformed coordinates of the matched atoms
    DevBuf o_off, o_fill, o_grouped, o_order, o_tiles, o_runs, o_work;   // ordering workspace
    DevBuf k_key, k_key2, k_iota, k_runid, k_scan;             // keyed hit ordering (hits.cuh)
    DevBuf h_pair, h_keys;                                     // reference-order mode: pair of each hit, rank vectors
    DevBuf q_ocnt, q_keyA, q_keyA2, q_keyB, q_keyB2, q_iota, q_perm, q_setpos, q_kdrank, q_coords, q_kdidx, q_cub;
    DevBuf q_need, q_pairhits;                                 // reference order: lists whose ranks are read, hits per pair
    RunWork *host_work = nullptr;                              // pinned
    std::vector<std::pair<void *, size_t>> pin_pool;           // page-locked result blocks, recycled
    std::mutex pool_mu;                                        // jb_result_free may run on another thread than a query
    long long hit_cap = 0;
    int hit_stride = 0;
    size_t decay_bytes = (size_t)1 << 30;   // JB_DECAY_BYTES
    int decay_runs = 64;                    // JB_DECAY_RUNS
    int runs_since_decay = 0;     // context_decay: every kDecayRuns calls, buffers far larger than recently needed go back
    long long trim_epoch = 0;     // bumped by jb_context_trim: un-fetched results of earlier runs are gone
    long long run_epoch = 0;      // bumped by every engine run: the hit buffers hold the records of the LAST run only
    Counters *host_ctr = nullptr; // pinned
    u64 *host_lv = nullptr;       // pinned, JB_DEBUG: (scan_tests, box_pass) after every level
    struct jb_molecules *scratch_mol = nullptr;   // reused by jb_query_host (no per-call cudaMalloc)
    struct jb_molecules *slice_mol[2] = { nullptr, nullptr };   // double-buffered staging of the pipelined jb_query_host
    cudaStream_t copy_stream = nullptr;           // H2D of the next slice / D2H of the previous slice's hits
    PinBuf host_stage[2];                         // jb_query_records: page-locked SoA image of the slice being packed
    cudaEvent_t ev_up[2] = { nullptr, nullptr };
    int pipeline_slices = 0;                      // JB_PIPELINE: 0 auto, 1 off, n >= 2 forced
    long long last_hits = 0;                      // sizes the pinned result block of the next pipelined call
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
};

struct jb_templates {
    jb_context *ctx = nullptr;
    int T = 0, TA = 0, S = 0, nmax = 0, nmin = 1 << 30;
    std::vector<int> atom_off;
    DevBuf d_atom_off, d_spec, d_resseq, d_chain, d_pos, d_weight, d_dist_off, d_dist, d_shell, d_vol;
    DevBuf s_code, s_noff, s_nraw, s_nup, s_roff, s_rup;
    int OS = 0;                                   // ordered specs (order.cuh)
    bool order_ok = true;                         // false: an atom lists > 31 distinct residue names
    DevBuf o_spec, o_byres, o_roff, o_res, d_ospec;
    long long n_dist = 0;
    // the shell radii (k_shells) of the last query: a service asks the same thresholds again and again
    mutable bool shells_valid = false;
    mutable double shells_cutoff = 0.0, shells_maxdyn = 0.0;
};

struct jb_molecules {
    jb_context *ctx = nullptr;
    int M = 0;
    long long NA = 0;
    std::vector<long long> atom_off;
    DevBuf d_off, x, y, z, name, res, resseq, chain, atoms, serial;
    bool has_serial = false;
    DevBuf srec, sname, sres, ext, sort_scratch;   // x-sorted FP32 view (k_sort_atoms)
    std::vector<DevBuf *> all() { return { &d_off, &x, &y, &z, &name, &res, &resseq, &chain, &atoms, &serial, &srec, &sname, &sres, &ext, &sort_scratch }; }
};

struct jb_result {
    jb_context *ctx = nullptr;
    long long count = 0;      // hits in the device buffers of ctx (valid until the next run)
    int stride = 0;
    int best_match = 0;
    int n_templates = 0;
    jb_stats stats{};
    bool fetched = false;
    long long trim_epoch = 0;
    long long run_epoch = 0;
    void *block = nullptr;    // page-locked host block holding the fetched arrays (from ctx->pin_pool)
    size_t block_bytes = 0;
    size_t block_cap = 0;     // records the block is laid out for
    long long kept = 0;       // records after best_match selection
    int32_t *mol = nullptr, *tpl = nullptr, *atoms = nullptr;
    double *rmsd = nullptr, *rot = nullptr, *c0 = nullptr, *c1 = nullptr;
    int coords = 0;           // JB_QUERY_COORDS: xyz holds count * stride * 3 transformed coordinates
    double *xyz = nullptr;
    bool direct = false;      // small result: the ordering kernel wrote the records straight into `block`
};

static void context_trim(jb_context *c);

// ---- library -----------------------------------------------------------------------
extern "C" int jb_abi_version(void) { return JB_ABI_VERSION; }
extern "C" const char *jb_last_error(void) { return g_err; }
extern "C" int jb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// Page-locked host memory for the staging arrays of jb_molecule_batch (portable across contexts): an upload from
// it is one asynchronous DMA, from pageable memory the driver stages it through its own bounce buffers.
extern "C" int jb_host_alloc(int64_t bytes, void **out)
{
    if (!out || bytes < 0) return fail(JB_ERR_INVALID, "jb_host_alloc: bad argument");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(JB_ERR_CUDA, "no CUDA device available; the engine has no CPU fallback");
    }
    CU(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable));
    return JB_OK;
}

extern "C" void jb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int jb_context_create(int device, void *stream, int64_t workspace_bytes, jb_context **out)
{
    if (!out) return fail(JB_ERR_INVALID, "jb_context_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(JB_ERR_CUDA, "no CUDA device available (%s); the engine has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(JB_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
    CU(cudaSetDevice(device));
    jb_context *c = new jb_context();
    c->device = device;
    if (stream) c->stream = (cudaStream_t)stream;
    else { CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    c->ws_limit = workspace_bytes > 0 ? (size_t)workspace_bytes : total_b / 8;
    if (const char *g = getenv("JB_GROUP")) {
        int v = atoi(g);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) c->group = v;
    }
    // one lane per partial match at the hot levels: the cell-grid sweep works per lane, and a batch of 32
    // partial matches halves the per-batch overheads (prefix sums, final partial drain)
    c->group_lv[1] = c->group_lv[2] = c->group_lv[3] = 1;
    if (const char *g = getenv("JB_GROUPS")) {
        int l = 1;
        for (const char *p = g; *p && l <= kMaxN; l++) {
            c->group_lv[l] = atoi(p);
            while (*p && *p != ',') p++;
            if (*p == ',') p++;
        }
    }
    if (const char *g = getenv("JB_DECAY_BYTES")) c->decay_bytes = (size_t)std::max(0ll, atoll(g));
    if (const char *g = getenv("JB_DECAY_RUNS")) c->decay_runs = std::max(1, atoi(g));
    if (const char *g = getenv("JB_DEBUG")) c->debug = atoi(g);
    if (const char *g = getenv("JB_LEGACY")) c->legacy = atoi(g);
    if (const char *g = getenv("JB_SWEEP")) c->sweep = atoi(g);
    if (const char *g = getenv("JB_DFS_PAIRS")) c->dfs_pairs = atoll(g);
    if (const char *g = getenv("JB_TAIL_ITEMS")) c->tail_items = atoll(g);
    if (const char *g = getenv("JB_FIRST_SLICE")) c->first_slice_pct = std::max(0, std::min(90, atoi(g)));
    if (const char *g = getenv("JB_GRID_LEVELS")) c->grid_levels = std::max(1, std::min(24, atoi(g)));
    if (const char *g = getenv("JB_PIPELINE")) c->pipeline_slices = atoi(g);
    if (c->debug) for (int i = 0; i <= kMaxN; i++) CU(cudaEventCreate(&c->lev[i]));
    CU(cudaMallocHost(&c->host_ctr, sizeof(Counters)));
    CU(cudaMallocHost(&c->host_work, sizeof(RunWork)));
    CU(cudaMallocHost(&c->host_lv, 2 * (kMaxN + 2) * sizeof(u64)));
    CU(c->o_work.ensure(2 * sizeof(RunWork), c->stream));   // [0] tile work list, [1].n_runs = number of tie runs
    CU(c->ctr.ensure(sizeof(Counters), c->stream));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    for (int i = 0; i < 5; i++) CU(cudaEventCreate(&c->evs[i]));
    *out = c;
    return JB_OK;
}

// Every grow-only device buffer of the context except the two small ones allocated at creation (counters,
// ordering work list).  They only ever grow to the largest workload seen; jb_context_trim gives the memory back.
static std::vector<DevBuf *> context_buffers(jb_context *c)
{
    return { &c->crec, &c->chead, &c->cxtab, &c->cbin, &c->cbinlist, &c->recs, &c->ordbuf, &c->fr[0], &c->fr[1],
             &c->tuples, &c->h_mol, &c->h_tpl, &c->h_atoms, &c->h_rmsd, &c->h_rot, &c->h_c0,
             &c->h_c1, &c->f_mol, &c->f_tpl, &c->f_atoms, &c->f_rmsd, &c->f_rot, &c->f_c0, &c->f_c1, &c->f_xyz,
             &c->o_off, &c->o_fill, &c->o_grouped, &c->o_order, &c->o_tiles, &c->o_runs,
             &c->k_key, &c->k_key2, &c->k_iota, &c->k_runid, &c->k_scan, &c->h_pair, &c->h_keys, &c->q_ocnt, &c->q_keyA,
             &c->q_keyA2, &c->q_keyB, &c->q_keyB2, &c->q_iota, &c->q_perm, &c->q_setpos, &c->q_kdrank, &c->q_coords,
             &c->q_kdidx, &c->q_cub, &c->q_need, &c->q_pairhits };
}

static void context_trim(jb_context *c)
{
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    cudaGetLastError();                               // a failed cudaMalloc leaves its (non-sticky) error behind
    for (DevBuf *b : context_buffers(c)) b->release();
    for (jb_molecules *m : { c->scratch_mol, c->slice_mol[0], c->slice_mol[1] })
        if (m) for (DevBuf *b : m->all()) b->release();
    for (PinBuf &b : c->host_stage) b.release();
    {
        std::lock_guard<std::mutex> g(c->pool_mu);
        for (auto &blk : c->pin_pool) cudaFreeHost(blk.first);
        c->pin_pool.clear();
    }
    c->hit_cap = 0;
    c->hit_stride = 0;
    c->trim_epoch++;                                  // results not fetched yet lived in these buffers
}

// High-water marks are not forever: every decay_runs (64) engine calls, a buffer of more than decay_bytes
// (1 GiB) that no call since the last check used an eighth of is released (one pathological query -- e.g.
// every tuple of a 3-atom template -- otherwise pins tens of GB for the life of the context).  Lazy on
// purpose: re-growing a frontier buffer costs a re-run of the chunk that overflowed, so a service that
// alternates batch sizes (or bench.py, whose sliced end-to-end calls use a quarter of what its resident
// calls sized) must not pay that every few calls -- measured with 8 calls / a quarter: cfg5 end to end
// 85 -> 153 ms.  Called at the start of a call: the previous run's un-fetched hits are invalid from here on.
static void context_decay(jb_context *c)
{
    if (++c->runs_since_decay < c->decay_runs) return;
    c->runs_since_decay = 0;
    std::vector<DevBuf *> bufs = context_buffers(c);
    for (jb_molecules *m : { c->scratch_mol, c->slice_mol[0], c->slice_mol[1] })
        if (m) for (DevBuf *b : m->all()) bufs.push_back(b);
    bool synced = false, released = false;
    for (DevBuf *b : bufs) {
        if (b->bytes > c->decay_bytes && b->recent < b->bytes / 8) {
            if (!synced) {
                cudaSetDevice(c->device);
                cudaStreamSynchronize(c->stream);
                if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
                synced = true;
            }
            b->release();
            released = true;
        }
        b->recent = 0;
    }
    if (released) {
        // hit_cap describes ALL hit arrays: it only survives when none of them went (a needless reset would cost
        // the next large run a chunk re-run through grow_hits)
        DevBuf *hb[] = { &c->h_mol, &c->h_tpl, &c->h_atoms, &c->h_rmsd, &c->h_rot, &c->h_c0, &c->h_c1, &c->f_mol, &c->f_tpl,
                         &c->f_atoms, &c->f_rmsd, &c->f_rot, &c->f_c0, &c->f_c1, &c->h_pair, &c->h_keys };
        for (DevBuf *b : hb)
            if (b->bytes == 0) { c->hit_cap = 0; c->hit_stride = 0; break; }
    }
}

extern "C" int jb_context_trim(jb_context *c)
{
    if (!c) return fail(JB_ERR_INVALID, "jb_context_trim: NULL context");
    context_trim(c);
    return JB_OK;
}

extern "C" void jb_context_destroy(jb_context *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    std::vector<DevBuf *> bufs = context_buffers(c);
    bufs.push_back(&c->ctr);
    bufs.push_back(&c->o_work);
    for (DevBuf *b : bufs) b->release();
    for (jb_molecules *m : { c->scratch_mol, c->slice_mol[0], c->slice_mol[1] })
        if (m) {
            for (DevBuf *b : m->all()) b->release();
            delete m;
        }
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
        for (int i = 0; i < 2; i++) if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]);
    }
    for (PinBuf &b : c->host_stage) b.release();
    if (c->host_ctr) cudaFreeHost(c->host_ctr);
    if (c->host_work) cudaFreeHost(c->host_work);
    if (c->host_lv) cudaFreeHost(c->host_lv);
    for (auto &blk : c->pin_pool) cudaFreeHost(blk.first);
    c->pin_pool.clear();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 5; i++) if (c->evs[i]) cudaEventDestroy(c->evs[i]);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int jb_context_synchronize(jb_context *c)
{
    if (!c) return fail(JB_ERR_INVALID, "NULL context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return JB_OK;
}

// ---- templates ------------------------------------------------------------------------
static uint32_t upper4_host(uint32_t t)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t c = (t >> (8 * i)) & 0xffu;
        if (c >= 'a' && c <= 'z') c &= ~0x20u;
        r |= c << (8 * i);
    }
    return r;
}

static bool mode_valid(int code) { return (code >= -1 && code <= 8) || (code >= 100 && code <= 107); }
static bool mode_uses_names(int code) { return code == -1 || code == 0 || code == 3 || code == 8 || code == 100 || code == 103; }
static bool mode_uses_resnames(int code) { return code < 100; }

extern "C" int jb_templates_create(jb_context *c, const jb_template_table *tb, jb_templates **out)
{
    if (!c || !tb || !out) return fail(JB_ERR_INVALID, "jb_templates_create: NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(c->device));
    const int T = tb->n_templates;
    if (T < 0) return fail(JB_ERR_INVALID, "negative template count");
    jb_templates *t = new jb_templates();
    t->ctx = c;
    t->T = T;
    t->atom_off.assign(tb->atom_offset, tb->atom_offset + T + 1);
    t->TA = T ? t->atom_off[T] : 0;
    for (int i = 0; i < T; i++) {
        const int n = t->atom_off[i + 1] - t->atom_off[i];
        if (n > JB_MAX_TEMPLATE_ATOMS) {
            delete t;
            return fail(JB_ERR_INVALID, "template %d has %d atoms; the GPU path supports at most %d", i, n,
                        JB_MAX_TEMPLATE_ATOMS);
        }
        t->nmax = std::max(t->nmax, n);
        t->nmin = std::min(t->nmin, n);
    }
    // de-duplicate atom specs: (mode, set of names it reads, set of residue names it reads)
    std::map<std::vector<uint32_t>, int> seen;
    std::vector<int> spec(t->TA), s_code, s_noff{0}, s_roff{0};
    std::vector<uint32_t> s_nraw, s_nup, s_rup;
    for (int a = 0; a < t->TA; a++) {
        const int code = tb->match_mode[a];
        if (!mode_valid(code)) { delete t; return fail(JB_ERR_INVALID, "template atom %d: invalid match mode %d", a, code); }
        std::vector<uint32_t> names, res;
        if (mode_uses_names(code))
            names.assign(tb->atom_names + tb->name_offset[a], tb->atom_names + tb->name_offset[a + 1]);
        if (mode_uses_resnames(code))
            for (int k = tb->resname_offset[a]; k < tb->resname_offset[a + 1]; k++) res.push_back(upper4_host(tb->residue_names[k]));
        std::sort(names.begin(), names.end()); names.erase(std::unique(names.begin(), names.end()), names.end());
        std::sort(res.begin(), res.end()); res.erase(std::unique(res.begin(), res.end()), res.end());
        std::vector<uint32_t> key{ (uint32_t)code, (uint32_t)names.size() };
        key.insert(key.end(), names.begin(), names.end());
        key.insert(key.end(), res.begin(), res.end());
        auto it = seen.find(key);
        if (it == seen.end()) {
            const int id = (int)s_code.size();
            seen.emplace(key, id);
            s_code.push_back(code);
            for (uint32_t v : names) { s_nraw.push_back(v); s_nup.push_back(upper4_host(v)); }
            for (uint32_t v : res) s_rup.push_back(v);
            s_noff.push_back((int)s_nraw.size());
            s_roff.push_back((int)s_rup.size());
            spec[a] = id;
        } else spec[a] = it->second;
    }
    t->S = (int)s_code.size();
    // ordered specs (order.cuh): engine spec + the residue names in the order written
    // (duplicates skipped like the `done` array of eff:TessTemplate.c:160-176)
    std::map<std::vector<uint32_t>, int> oseen;
    std::vector<int> ospec(t->TA), o_spec, o_byres, o_roff{0};
    std::vector<uint32_t> o_res;
    for (int a = 0; a < t->TA; a++) {
        const int code = tb->match_mode[a];
        const bool by_res = code >= -1 && code <= 8;
        std::vector<uint32_t> ord;
        if (by_res)
            for (int k = tb->resname_offset[a]; k < tb->resname_offset[a + 1]; k++) {
                const uint32_t v = upper4_host(tb->residue_names[k]);
                if (std::find(ord.begin(), ord.end(), v) == ord.end()) ord.push_back(v);
            }
        if (ord.size() > 31) t->order_ok = false;
        std::vector<uint32_t> key{ (uint32_t)spec[a] };
        key.insert(key.end(), ord.begin(), ord.end());
        auto it = oseen.find(key);
        if (it == oseen.end()) {
            const int id = (int)o_spec.size();
            oseen.emplace(key, id);
            o_spec.push_back(spec[a]);
            o_byres.push_back(by_res ? 1 : 0);
            for (uint32_t v : ord) o_res.push_back(v);
            o_roff.push_back((int)o_res.size());
            ospec[a] = id;
        } else ospec[a] = it->second;
    }
    t->OS = (int)o_spec.size();
    t->n_dist = T ? tb->distance_offset[T] : 0;
    cudaStream_t st = c->stream;
    std::vector<long long> dist_off(tb->distance_offset, tb->distance_offset + T + 1);
    cudaError_t e = cudaSuccess;
#define UP(buf, ptr, n) if (e == cudaSuccess) e = upload(buf, ptr, (size_t)(n), st)
    UP(t->d_atom_off, t->atom_off.data(), T + 1);
    UP(t->d_spec, spec.data(), t->TA);
    UP(t->d_resseq, tb->residue_number, t->TA);
    UP(t->d_chain, tb->chain_id2, t->TA);
    UP(t->d_pos, tb->position, 3 * (size_t)t->TA);
    UP(t->d_weight, tb->distance_weight, t->TA);
    UP(t->d_dist_off, dist_off.data(), T + 1);
    UP(t->d_dist, tb->distances, t->n_dist);
    UP(t->s_code, s_code.data(), s_code.size());
    UP(t->s_noff, s_noff.data(), s_noff.size());
    UP(t->s_nraw, s_nraw.data(), s_nraw.size());
    UP(t->s_nup, s_nup.data(), s_nup.size());
    UP(t->s_roff, s_roff.data(), s_roff.size());
    UP(t->s_rup, s_rup.data(), s_rup.size());
    UP(t->o_spec, o_spec.data(), o_spec.size());
    UP(t->o_byres, o_byres.data(), o_byres.size());
    UP(t->o_roff, o_roff.data(), o_roff.size());
    UP(t->o_res, o_res.data(), o_res.size());
    UP(t->d_ospec, ospec.data(), ospec.size());
#undef UP
    if (e == cudaSuccess) e = t->d_shell.ensure(std::max<size_t>(t->n_dist, 1) * sizeof(ShellRec), st);
    if (e == cudaSuccess) e = t->d_vol.ensure(std::max<size_t>(t->n_dist, 1) * sizeof(double), st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // host vectors die at return
    if (e != cudaSuccess) {
        jb_templates_free(t);
        return fail(e == cudaErrorMemoryAllocation ? JB_ERR_NOMEM : JB_ERR_CUDA, "template upload: %s", cudaGetErrorString(e));
    }
    *out = t;
    return JB_OK;
}

extern "C" void jb_templates_free(jb_templates *t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    DevBuf *bufs[] = { &t->d_atom_off, &t->d_spec, &t->d_resseq, &t->d_chain, &t->d_pos, &t->d_weight, &t->d_dist_off,
                       &t->d_dist, &t->d_shell, &t->d_vol, &t->s_code, &t->s_noff, &t->s_nraw, &t->s_nup, &t->s_roff,
                       &t->s_rup, &t->o_spec, &t->o_byres, &t->o_roff, &t->o_res, &t->d_ospec };
    for (DevBuf *b : bufs) b->release();
    delete t;
}

extern "C" int jb_templates_spec_count(const jb_templates *t) { return t ? t->S : 0; }

// ---- molecules --------------------------------------------------------------------------
// Host -> device copies of structures [m0, m1) of `b` on stream `st` (no kernels, no sync)
static int stage_molecules(jb_context *c, jb_molecules *m, const jb_molecule_batch *b, int m0, int m1, cudaStream_t st)
{
    const int M = m1 - m0;
    if (M < 0 || m0 < 0 || m1 > b->n_molecules) return fail(JB_ERR_INVALID, "bad molecule range [%d, %d)", m0, m1);
    m->ctx = c;
    m->M = M;
    m->atom_off.resize((size_t)M + 1);
    const long long a0 = b->atom_offset[m0];
    for (int i = 0; i <= M; i++) m->atom_off[i] = b->atom_offset[m0 + i] - a0;
    m->NA = M ? m->atom_off[M] : 0;
    for (int i = 0; i < M; i++)
        if (m->atom_off[i + 1] < m->atom_off[i] || m->atom_off[i + 1] - m->atom_off[i] > 0x7fffffffLL)
            return fail(JB_ERR_INVALID, "molecule %d: bad atom offsets", m0 + i);
    cudaError_t e = cudaSuccess;
#define UP(buf, ptr, n) if (e == cudaSuccess) e = upload(buf, ptr, (size_t)(n), st)
    UP(m->d_off, m->atom_off.data(), M + 1);
    UP(m->x, b->x + a0, m->NA);
    UP(m->y, b->y + a0, m->NA);
    UP(m->z, b->z + a0, m->NA);
    UP(m->name, b->name + a0, m->NA);
    UP(m->res, b->residue_name + a0, m->NA);
    UP(m->resseq, b->residue_number + a0, m->NA);
    UP(m->chain, b->chain_id2 + a0, m->NA);
    m->has_serial = b->serial != nullptr;
    if (m->has_serial) UP(m->serial, b->serial + a0, m->NA);
#undef UP
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? JB_ERR_NOMEM : JB_ERR_CUDA, "molecule upload: %s", cudaGetErrorString(e));
    return JB_OK;
}

// Derived device images of staged structures: 32-byte atom records, per-structure frame, x-sorted view
static int prepare_molecules(jb_context *c, jb_molecules *m, cudaStream_t st)
{
    (void)c;
    const int M = m->M;
    // the sweep addresses atoms and candidate records with 32-bit offsets (expand_grid.cuh: GridWarp)
    if (m->NA >= (1ll << 32)) return fail(JB_ERR_INVALID, "an upload holds %lld atoms; the engine takes fewer than 2^32 per upload", m->NA);
    cudaError_t e = m->atoms.ensure(std::max<size_t>((size_t)m->NA, 1) * sizeof(AtomRec), st);
    long long amax_all = 0;
    for (int i = 0; i < M; i++) amax_all = std::max(amax_all, m->atom_off[i + 1] - m->atom_off[i]);
    // small batches: the sort kernel builds the 32-byte records of its structure itself (one launch less)
    const bool fused_pack = M > 0 && M <= 64 && amax_all <= 16384;
    const bool radix_small = fused_pack && amax_all <= 8 * kPrepThreads && !getenv("JB_NO_SMALL_PREPARE");
    if (e == cudaSuccess && m->NA > 0 && !fused_pack) {
        k_pack_atoms<<<(unsigned)((m->NA + 255) / 256), 256, 0, st>>>(m->NA, m->x.as<double>(), m->y.as<double>(),
                                                                     m->z.as<double>(), m->resseq.as<int>(),
                                                                     m->chain.as<unsigned char>(), m->atoms.as<AtomRec>());
        e = cudaGetLastError();
    }
    // per-structure frame + x-sorted FP32 view (replaces the reference's per-pair k-d tree builds)
    if (e == cudaSuccess) e = m->srec.ensure(std::max<size_t>((size_t)m->NA, 1) * sizeof(float4), st);
    if (e == cudaSuccess) e = m->sname.ensure(std::max<size_t>((size_t)m->NA, 1) * sizeof(uint32_t), st);
    if (e == cudaSuccess) e = m->sres.ensure(std::max<size_t>((size_t)m->NA, 1) * sizeof(uint32_t), st);
    if (e == cudaSuccess) e = m->ext.ensure(std::max<size_t>((size_t)M, 1) * sizeof(MolExt), st);
    if (e == cudaSuccess && radix_small) {
        // latency path: records + frame + x-sorted view by one block radix sort per structure (prepare_small.cuh)
        const RawAtoms raw{ m->x.as<double>(), m->y.as<double>(), m->z.as<double>(), m->resseq.as<int>(), m->chain.as<unsigned char>() };
        if (amax_all <= 4 * kPrepThreads)
            k_prepare_small<4><<<M, kPrepThreads, 0, st>>>(m->d_off.as<long long>(), m->atoms.as<AtomRec>(), m->name.as<uint32_t>(),
                                                          m->res.as<uint32_t>(), m->ext.as<MolExt>(), m->srec.as<float4>(),
                                                          m->sname.as<uint32_t>(), m->sres.as<uint32_t>(), raw);
        else
            k_prepare_small<8><<<M, kPrepThreads, 0, st>>>(m->d_off.as<long long>(), m->atoms.as<AtomRec>(), m->name.as<uint32_t>(),
                                                          m->res.as<uint32_t>(), m->ext.as<MolExt>(), m->srec.as<float4>(),
                                                          m->sname.as<uint32_t>(), m->sres.as<uint32_t>(), raw);
        e = cudaGetLastError();
    } else if (e == cudaSuccess && M > 0) {
        long long amax = 0;
        for (int i = 0; i < M; i++) amax = std::max(amax, m->atom_off[i + 1] - m->atom_off[i]);
        if (amax > 16384 && m->NA < (1ll << 32)) {
            // a very large structure in the batch: one device-wide radix sort (kernels.cuh, K0c)
            const size_t n = (size_t)m->NA;
            e = m->sort_scratch.ensure(n * (2 * sizeof(u64) + 2 * sizeof(unsigned)), st);
            u64 *key = m->sort_scratch.as<u64>(), *key2 = key + n;
            unsigned *idx = reinterpret_cast<unsigned *>(key2 + n), *idx2 = idx + n;
            int mbits = 1;
            while ((1ll << mbits) < M) mbits++;
            size_t tmp = 0;
            if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tmp, key, key2, idx, idx2, (int)n, 0, 32 + mbits, st);
            if (e == cudaSuccess) e = c->q_cub.ensure(tmp, st);
            if (e == cudaSuccess) {
                k_sort_keys<<<M, kSortThreads, 0, st>>>(m->d_off.as<long long>(), m->atoms.as<AtomRec>(), m->ext.as<MolExt>(), key, idx);
                e = cub::DeviceRadixSort::SortPairs(c->q_cub.p, tmp, key, key2, idx, idx2, (int)n, 0, 32 + mbits, st);
            }
            if (e == cudaSuccess) {
                k_sorted_view<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((long long)n, key2, idx2, m->d_off.as<long long>(),
                                                                           m->atoms.as<AtomRec>(), m->name.as<uint32_t>(),
                                                                           m->res.as<uint32_t>(), m->ext.as<MolExt>(),
                                                                           m->srec.as<float4>(), m->sname.as<uint32_t>(),
                                                                           m->sres.as<uint32_t>());
                e = cudaGetLastError();
            }
        } else {
        long long npad = 1;
        while (npad < amax) npad <<= 1;
        const int smem_keys = (int)std::min<long long>(npad, 16384);      // 8-byte keys: <= 128 KiB
        if (npad > smem_keys) e = m->sort_scratch.ensure((size_t)m->NA * 2 * sizeof(u64), st);
        const size_t smem = (size_t)smem_keys * sizeof(u64);
        if (e == cudaSuccess && smem > 48 * 1024)
            e = cudaFuncSetAttribute(k_sort_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            RawAtoms raw{ nullptr, nullptr, nullptr, nullptr, nullptr };
            if (fused_pack) raw = RawAtoms{ m->x.as<double>(), m->y.as<double>(), m->z.as<double>(), m->resseq.as<int>(),
                                            m->chain.as<unsigned char>() };
            k_sort_atoms<<<M, kSortThreads, smem, st>>>(m->d_off.as<long long>(), m->atoms.as<AtomRec>(),
                                                        m->name.as<uint32_t>(), m->res.as<uint32_t>(),
                                                        m->ext.as<MolExt>(), m->srec.as<float4>(),
                                                        m->sname.as<uint32_t>(), m->sres.as<uint32_t>(),
                                                        m->sort_scratch.as<u64>(), smem_keys, raw);
            e = cudaGetLastError();
        }
        }
    }
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? JB_ERR_NOMEM : JB_ERR_CUDA, "molecule upload: %s", cudaGetErrorString(e));
    return JB_OK;
}

// Where the structures of a host call come from: a SoA image packed by the caller (jb_query_host), or the host
// layer's atom records, which the engine packs itself slice by slice (jb_query_records).
struct HostSource {
    const jb_molecule_batch *soa = nullptr;
    const jb_atom_record *const *recs = nullptr;
    const int64_t *counts = nullptr;
    int M = 0;
    bool serial = false;
    std::vector<int64_t> off;                         // [M + 1] atom offsets (records mode)
    const int64_t *offsets() const { return soa ? soa->atom_offset : off.data(); }
};

// Structures [m0, m1) of the source -> device staging `m` on stream `st` (copies only, no kernels, no sync).
// Records mode packs them first into the page-locked block `stg` (jb_molecules_pack: one host thread per share of the
// atoms), which must stay untouched until the stream has passed the copies.
static int stage_source(jb_context *c, jb_molecules *m, const HostSource &src, int m0, int m1, cudaStream_t st, PinBuf &stg)
{
    if (src.soa) return stage_molecules(c, m, src.soa, m0, m1, st);
    const int M = m1 - m0;
    const int64_t NA = src.off[m1] - src.off[m0];
    const size_t n = (size_t)std::max<int64_t>(NA, 1);
    auto up8 = [](size_t v) { return (v + 7) & ~(size_t)7; };
    const size_t o_x = 0, o_y = o_x + 8 * n, o_z = o_y + 8 * n, o_name = o_z + 8 * n, o_res = o_name + up8(4 * n),
                 o_seq = o_res + up8(4 * n), o_ser = o_seq + up8(4 * n), o_chain = o_ser + (src.serial ? up8(4 * n) : 0),
                 o_off = o_chain + up8(n), total = o_off + 8 * ((size_t)M + 1);
    if (stg.ensure(total) != cudaSuccess) { cudaGetLastError(); return fail(JB_ERR_NOMEM, "page-locked staging of %zu bytes", total); }
    char *base = stg.as<char>();
    jb_molecule_batch view;
    view.n_molecules = M;
    int64_t *off = reinterpret_cast<int64_t *>(base + o_off);
    double *x = reinterpret_cast<double *>(base + o_x), *y = reinterpret_cast<double *>(base + o_y), *z = reinterpret_cast<double *>(base + o_z);
    uint32_t *name = reinterpret_cast<uint32_t *>(base + o_name), *res = reinterpret_cast<uint32_t *>(base + o_res);
    int32_t *seq = reinterpret_cast<int32_t *>(base + o_seq), *ser = src.serial ? reinterpret_cast<int32_t *>(base + o_ser) : nullptr;
    uint8_t *chain = reinterpret_cast<uint8_t *>(base + o_chain);
    const int rc = jb_molecules_pack(src.recs + m0, src.counts + m0, M, off, x, y, z, name, res, seq, chain, ser);
    if (rc != JB_OK) return fail(rc, "jb_molecules_pack failed on structures [%d, %d)", m0, m1);
    view.atom_offset = off; view.x = x; view.y = y; view.z = z; view.name = name; view.residue_name = res;
    view.residue_number = seq; view.chain_id2 = chain; view.serial = ser;
    return stage_molecules(c, m, &view, 0, M, st);
}

static int fill_molecules(jb_context *c, jb_molecules *m, const jb_molecule_batch *b)
{
    const double t_0 = now_ms();
    if (b->n_molecules < 0) return fail(JB_ERR_INVALID, "negative molecule count");
    int rc = stage_molecules(c, m, b, 0, b->n_molecules, c->stream);
    if (rc == JB_OK) rc = prepare_molecules(c, m, c->stream);
    if (rc != JB_OK) return rc;
    // the caller's staging buffers must stay valid until the stream reaches the copies:
    // synchronise so that they may be reused immediately.
    CU(cudaStreamSynchronize(c->stream));
    if (c->debug) fprintf(stderr, "[jb] upload %d structures, %lld atoms: %.2f ms\n", m->M, m->NA, now_ms() - t_0);
    return JB_OK;
}

extern "C" int jb_molecules_upload(jb_context *c, const jb_molecule_batch *b, jb_molecules **out)
{
    if (!c || !b || !out) return fail(JB_ERR_INVALID, "jb_molecules_upload: NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(c->device));
    jb_molecules *m = new jb_molecules();
    m->ctx = c;
    int rc = fill_molecules(c, m, b);
    if (rc == JB_ERR_NOMEM) {                         // the context's idle buffers may hold the memory (see jb_query_run)
        context_trim(c);
        rc = fill_molecules(c, m, b);
    }
    if (rc != JB_OK) { jb_molecules_free(m); return rc; }
    *out = m;
    return JB_OK;
}

extern "C" void jb_molecules_free(jb_molecules *m)
{
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (DevBuf *b : m->all()) b->release();
    delete m;
}

// ---- query --------------------------------------------------------------------------------
static size_t expand_smem(int group, int level) { return (size_t)kExpandWarps * expand_warp_bytes(group, level); }

template <int G, bool STATS> static cudaError_t launch_expand(int grid, size_t smem, cudaStream_t st, const ExpandArgs &a)
{
    if (smem > 48 * 1024) {   // per device: set on every launch (a context per GPU may share this process)
        cudaError_t e = cudaFuncSetAttribute(k_expand<G, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_expand<G, STATS><<<grid, kExpandThreads, smem, st>>>(a);
    return cudaGetLastError();
}

// level-specialised kernel (expand_lv.cuh): levels 1..kMaxLv, G lanes per partial match
template <int L, int G, bool ROOT = false>
static cudaError_t launch_expand_lv(int sm_count, cudaStream_t st, const ExpandArgs &a, int stats)
{
    constexpr size_t smem = expand_lv_smem<L, G>();
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_expand_lv<L, G, ROOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / smem));
    k_expand_lv<L, G, ROOT><<<sm_count * per_sm, kExpandThreads, smem, st>>>(a, stats);
    return cudaGetLastError();
}

static cudaError_t dispatch_expand_lv(int level, int g, int sm_count, cudaStream_t st, const ExpandArgs &a, int stats)
{
    constexpr size_t budget = 200 * 1024;   // keep at least two CTAs per SM resident
#define JB_LV_CASE(LL)                                                                                             \
    case LL:                                                                                                       \
        if (g <= 1 && 2 * expand_lv_smem<LL, 1>() <= budget) return launch_expand_lv<LL, 1>(sm_count, st, a, stats); \
        if (g <= 2 && 2 * expand_lv_smem<LL, 2>() <= budget) return launch_expand_lv<LL, 2>(sm_count, st, a, stats); \
        return launch_expand_lv<LL, 4>(sm_count, st, a, stats);
    switch (level) {
        JB_LV_CASE(1) JB_LV_CASE(2) JB_LV_CASE(3) JB_LV_CASE(4)
        JB_LV_CASE(5) JB_LV_CASE(6) JB_LV_CASE(7) JB_LV_CASE(8)
        JB_LV_CASE(9) JB_LV_CASE(10) JB_LV_CASE(11) JB_LV_CASE(12)
    default: return cudaErrorInvalidValue;
    }
#undef JB_LV_CASE
}

// cell-grid sweep (expand_grid.cuh): levels 1..kMaxLv, G lanes per partial match
template <int L, int G, bool ROOT = false>
static cudaError_t launch_expand_grid(int sm_count, cudaStream_t st, const ExpandArgs &a, const GridArgs &b, int stats)
{
    constexpr size_t smem = expand_grid_smem<L, G>();
    if (smem > 48 * 1024) {   // per device: set on every launch (a context per GPU may share this process)
        cudaError_t e = cudaFuncSetAttribute(k_expand_grid<L, G, ROOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (228 * 1024) / (smem + 1024)));   // 228 KB per SM, 1 KB reserved per CTA
    k_expand_grid<L, G, ROOT><<<sm_count * per_sm, kExpandThreads, smem, st>>>(a, b, stats);
    return cudaGetLastError();
}

static cudaError_t dispatch_expand_grid(int level, int g, int sm_count, cudaStream_t st, const ExpandArgs &a, const GridArgs &b,
                                        int stats)
{
    constexpr size_t budget = 200 * 1024;   // keep at least two CTAs per SM resident
#define JB_GRID_CASE(LL)                                                                                                  \
    case LL:                                                                                                              \
        if (g <= 1 && 2 * expand_grid_smem<LL, 1>() <= budget) return launch_expand_grid<LL, 1>(sm_count, st, a, b, stats); \
        if (g <= 2 && 2 * expand_grid_smem<LL, 2>() <= budget) return launch_expand_grid<LL, 2>(sm_count, st, a, b, stats); \
        return launch_expand_grid<LL, 4>(sm_count, st, a, b, stats);
    switch (level) {
        JB_GRID_CASE(1) JB_GRID_CASE(2) JB_GRID_CASE(3) JB_GRID_CASE(4)
        JB_GRID_CASE(5) JB_GRID_CASE(6) JB_GRID_CASE(7) JB_GRID_CASE(8)
        JB_GRID_CASE(9) JB_GRID_CASE(10) JB_GRID_CASE(11) JB_GRID_CASE(12)
        // levels 13 .. kMaxGridLv (templates of up to 25 atoms, BASELINE config 5 holds 24): four lanes per partial match
        // -- the staging of 13-24 placed atoms for 16 or 32 partial matches per warp does not fit two CTAs per SM
#define JB_GRID_DEEP(LL) case LL: return launch_expand_grid<LL, 4>(sm_count, st, a, b, stats);
        JB_GRID_DEEP(13) JB_GRID_DEEP(14) JB_GRID_DEEP(15) JB_GRID_DEEP(16) JB_GRID_DEEP(17) JB_GRID_DEEP(18)
        JB_GRID_DEEP(19) JB_GRID_DEEP(20) JB_GRID_DEEP(21) JB_GRID_DEEP(22) JB_GRID_DEEP(23) JB_GRID_DEEP(24)
#undef JB_GRID_DEEP
    default: return cudaErrorInvalidValue;
    }
#undef JB_GRID_CASE
}

// Reference iteration order (order.cuh): builds the (structure, ordered spec) rank tables of
// the chunk and the rank vector of every hit in [h0, h1).
static int build_reference_keys(jb_context *c, const jb_templates *t, const jb_molecules *mol, const jb_query_params *q,
                                int m0, int Mc, const MolTable &MT, const CandTable &CT, const TemplateTable &TT,
                                long long h0, long long h1, jb_stats &S)
{
    cudaStream_t st = c->stream;
    if (!t->order_ok) return fail(JB_ERR_INVALID, "reference hit order: a template atom lists more than 31 residue names");
    const long long nl = (long long)Mc * t->OS;
    if (nl >= (1ll << kListBits)) return fail(JB_ERR_INVALID, "reference hit order: %lld candidate lists in one chunk", nl);
    OrderSpecs O;
    O.n_ospecs = t->OS; O.spec = t->o_spec.as<int>(); O.by_res = t->o_byres.as<int>(); O.res_off = t->o_roff.as<int>();
    O.res_up = t->o_res.as<uint32_t>(); O.atom_ospec = t->d_ospec.as<int>();
    const int n_tiles = (int)((nl + 1 + kScanTile - 1) / kScanTile);
    CU(c->q_ocnt.ensure((size_t)(nl + 1) * sizeof(unsigned), st));
    CU(c->o_tiles.ensure((size_t)n_tiles * sizeof(unsigned), st));
    unsigned *ooff = c->q_ocnt.as<unsigned>();
    // only the lists of pairs holding two or more hits (order.cuh)
    const long long n_pairs = (long long)Mc * t->T;
    CU(c->q_need.ensure((size_t)(nl + 1), st));
    CU(c->q_pairhits.ensure((size_t)n_pairs * sizeof(unsigned), st));
    CU(cudaMemsetAsync(c->q_need.p, 0, (size_t)(nl + 1), st));
    CU(cudaMemsetAsync(c->q_pairhits.p, 0, (size_t)n_pairs * sizeof(unsigned), st));
    {
        const unsigned hb = (unsigned)((h1 - h0 + 255) / 256);
        k_pair_hits<<<hb, 256, 0, st>>>(h0, h1, c->h_pair.as<int>(), c->q_pairhits.as<unsigned>());
        k_mark_lists<<<hb, 256, 0, st>>>(h0, h1, c->h_pair.as<int>(), c->q_pairhits.as<unsigned>(), c->recs.as<PairRec>(), O,
                                         c->q_need.as<unsigned char>());
        S.launches += 2;
    }
    k_olist_count<<<(unsigned)((nl + 1 + 255) / 256), 256, 0, st>>>(Mc, O, t->S, CT.head, c->q_need.as<unsigned char>(), ooff);
    k_scan_tiles<<<n_tiles, kScanThreads, 0, st>>>(ooff, nl + 1, c->o_tiles.as<unsigned>());
    k_scan_totals<<<1, kScanThreads, 0, st>>>(c->o_tiles.as<unsigned>(), n_tiles);
    k_scan_add<<<n_tiles, kScanThreads, 0, st>>>(ooff, nl + 1, c->o_tiles.as<unsigned>());
    CU(cudaGetLastError());
    unsigned total = 0;
    CU(cudaMemcpyAsync(&total, ooff + nl, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    S.launches += 4;
    const size_t n = total;
    if (n > 0) {
        CU(c->q_keyA.ensure(n * sizeof(u64), st));  CU(c->q_keyA2.ensure(n * sizeof(u64), st));
        CU(c->q_keyB.ensure(n * sizeof(u64), st));  CU(c->q_keyB2.ensure(n * sizeof(u64), st));
        CU(c->q_iota.ensure(n * sizeof(unsigned), st)); CU(c->q_perm.ensure(n * sizeof(unsigned), st));
        CU(c->q_setpos.ensure(n * sizeof(unsigned), st)); CU(c->q_kdrank.ensure(n * sizeof(unsigned), st));
        CU(c->q_coords.ensure(n * 3 * sizeof(double), st)); CU(c->q_kdidx.ensure(n * sizeof(int), st));
        k_olist_fill<<<(unsigned)((nl * 32 + 255) / 256), 256, 0, st>>>(
            Mc, O, t->S, CT.head, CT.rec, MT.atom_off, mol->res.as<uint32_t>(),
            mol->has_serial ? mol->serial.as<int>() : nullptr, ooff, c->q_need.as<unsigned char>(), c->q_keyA.as<u64>(),
            c->q_keyB.as<u64>());
        CU(cudaGetLastError());
        int list_bits = 1;
        while ((1ll << list_bits) < nl) list_bits++;
        // sort 1: (list, atom) ascending; carries the CandidateSet key
        size_t tmp1 = 0, tmp2 = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp1, c->q_keyA.as<u64>(), c->q_keyA2.as<u64>(), c->q_keyB.as<u64>(),
                                           c->q_keyB2.as<u64>(), (int)n, 0, 32 + list_bits, st));
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, c->q_keyB2.as<u64>(), c->q_keyB.as<u64>(), c->q_iota.as<unsigned>(),
                                           c->q_perm.as<unsigned>(), (int)n, 0, 37 + list_bits, st));
        CU(c->q_cub.ensure(std::max(tmp1, tmp2), st));
        CU(cub::DeviceRadixSort::SortPairs(c->q_cub.p, tmp1, c->q_keyA.as<u64>(), c->q_keyA2.as<u64>(), c->q_keyB.as<u64>(),
                                           c->q_keyB2.as<u64>(), (int)n, 0, 32 + list_bits, st));
        k_iota<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->q_iota.as<unsigned>(), (long long)n);
        // sort 2 (stable): by (list, residue-name position, serial); ties keep atom order
        CU(cub::DeviceRadixSort::SortPairs(c->q_cub.p, tmp2, c->q_keyB2.as<u64>(), c->q_keyB.as<u64>(), c->q_iota.as<unsigned>(),
                                           c->q_perm.as<unsigned>(), (int)n, 0, 37 + list_bits, st));
        k_olist_ranks<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            (long long)n, t->OS, c->q_keyB.as<u64>(), c->q_perm.as<unsigned>(), c->q_keyA2.as<u64>(), ooff, MT.atom_off,
            MT.atoms, c->q_setpos.as<unsigned>(), c->q_coords.as<double>(), c->q_kdidx.as<int>());
        // one thread per list, lists in memory order (a length-sorted schedule was measured 40 % slower:
        // the threads of a warp then walk lists of different structures and lose their shared cache lines)
        // lists of up to 32 records: one thread each; longer ones: one warp each, index array in shared memory
        // (order.cuh: k_kd_order_team, three size classes); beyond 16,384 records the thread kernel again (global memory,
        // slow but complete)
        constexpr int kTeamMid = 256, kTeamLong = 1024, kTeamXl = 4096, kTeamHuge = 16384;
        const bool team = !getenv("JB_NO_KD_TEAM");
        // upper list sizes of the thread / mid / long / xl kernels (the capacities above are their limits); tests lower them
        // with JB_KD_CLASSES=a,b,c,d so that small inputs drive every size class
        int lim[4] = { kKdThreadMax, kTeamMid, kTeamLong, kTeamXl };
        if (const char *g = getenv("JB_KD_CLASSES")) {
            int v[4];
            if (sscanf(g, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4 && v[0] >= 1 && v[0] <= v[1] && v[1] <= v[2] && v[2] <= v[3] &&
                v[1] <= kTeamMid && v[2] <= kTeamLong && v[3] <= kTeamXl)
                for (int i = 0; i < 4; i++) lim[i] = v[i];
        }
        k_kd_order<<<(unsigned)((nl + 127) / 128), 128, 0, st>>>(nl, nullptr, ooff, c->q_coords.as<double>(), c->q_kdidx.as<int>(),
                                                                 c->q_perm.as<unsigned>(), c->q_kdrank.as<unsigned>(),
                                                                 team ? (unsigned)lim[0] : ~0u, team ? (unsigned)kTeamHuge : ~0u);
        CU(cudaGetLastError());
        if (team) {
            // shared memory per warp: 8 bytes per record of the class's capacity (index array + two segment queues)
            constexpr size_t smem_mid = 8 * 2 * kTeamMid * sizeof(int), smem_long = 8 * 2 * kTeamLong * sizeof(int),
                             smem_xl = 4 * 2 * kTeamXl * sizeof(int), smem_huge = 2 * kTeamHuge * sizeof(int);
            const double *co = c->q_coords.as<double>();
            const unsigned *pm = c->q_perm.as<unsigned>();
            unsigned *kr = c->q_kdrank.as<unsigned>();
            k_kd_order_team<kTeamMid, 8><<<(unsigned)std::min<long long>((nl + 7) / 8, (long long)c->sm_count * 8), 256, smem_mid, st>>>(
                nl, lim[0], lim[1], ooff, co, pm, kr);
            CU(cudaGetLastError());
            CU(cudaFuncSetAttribute(k_kd_order_team<kTeamLong, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_long));
            k_kd_order_team<kTeamLong, 8><<<(unsigned)std::min<long long>((nl + 7) / 8, (long long)c->sm_count * 3), 256, smem_long, st>>>(
                nl, lim[1], lim[2], ooff, co, pm, kr);
            CU(cudaGetLastError());
            CU(cudaFuncSetAttribute(k_kd_order_team<kTeamXl, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_xl));
            k_kd_order_team<kTeamXl, 4><<<(unsigned)std::min<long long>((nl + 3) / 4, c->sm_count), 128, smem_xl, st>>>(
                nl, lim[2], lim[3], ooff, co, pm, kr);
            CU(cudaGetLastError());
            CU(cudaFuncSetAttribute(k_kd_order_team<kTeamHuge, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_huge));
            k_kd_order_team<kTeamHuge, 1><<<(unsigned)std::min<long long>(nl, c->sm_count), 32, smem_huge, st>>>(
                nl, lim[3], kTeamHuge, ooff, co, pm, kr);
            CU(cudaGetLastError());
            S.launches += 4;
        }
        S.launches += 6;   // + the radix sort passes (library kernels)
    }
    HitKeyArgs K;
    K.h0 = h0; K.h1 = h1; K.stride = c->hit_stride; K.n_specs = t->S;
    K.ref_reorder = q->hit_order == JB_ORDER_REFERENCE ? 1 : 0;
    K.pair = c->h_pair.as<int>(); K.atoms = c->h_atoms.as<int>(); K.recs = c->recs.as<PairRec>(); K.tpl_spec = TT.spec;
    K.head = CT.head; K.O = O; K.ooff = ooff; K.atom_sorted = c->q_keyA2.as<u64>();
    K.set_pos = c->q_setpos.as<unsigned>(); K.kd_rank = c->q_kdrank.as<unsigned>(); K.keys = c->h_keys.as<int>();
    k_hit_keys<<<(unsigned)((h1 - h0 + 127) / 128), 128, 0, st>>>(K);
    CU(cudaGetLastError());
    S.launches++;
    (void)m0;
    return JB_OK;
}

struct Range { int m0, m1; };

static int grow_hits(jb_context *c, long long need, int stride)
{
    if (need <= c->hit_cap && stride == c->hit_stride) return JB_OK;
    const long long cap = std::max<long long>(need + need / 4, 1 << 16);
    const size_t keep = (size_t)std::min(c->hit_cap, need);
    cudaStream_t st = c->stream;
    const bool same = stride == c->hit_stride;
    CU(c->h_pair.ensure(cap * sizeof(int), st, same ? keep * sizeof(int) : 0));
    CU(c->h_keys.ensure(cap * sizeof(int) * stride, st, 0));
    CU(c->h_mol.ensure(cap * sizeof(int), st, same ? keep * sizeof(int) : 0));
    CU(c->h_tpl.ensure(cap * sizeof(int), st, same ? keep * sizeof(int) : 0));
    CU(c->h_atoms.ensure(cap * sizeof(int) * stride, st, same ? keep * sizeof(int) * stride : 0));
    CU(c->h_rmsd.ensure(cap * sizeof(double), st, same ? keep * sizeof(double) : 0));
    CU(c->h_rot.ensure(cap * 9 * sizeof(double), st, same ? keep * 9 * sizeof(double) : 0));
    CU(c->h_c0.ensure(cap * 3 * sizeof(double), st, same ? keep * 3 * sizeof(double) : 0));
    CU(c->h_c1.ensure(cap * 3 * sizeof(double), st, same ? keep * 3 * sizeof(double) : 0));
    CU(c->f_mol.ensure(cap * sizeof(int), st, same ? keep * sizeof(int) : 0));
    CU(c->f_tpl.ensure(cap * sizeof(int), st, same ? keep * sizeof(int) : 0));
    CU(c->f_atoms.ensure(cap * sizeof(int) * stride, st, same ? keep * sizeof(int) * stride : 0));
    CU(c->f_rmsd.ensure(cap * sizeof(double), st, same ? keep * sizeof(double) : 0));
    CU(c->f_rot.ensure(cap * 9 * sizeof(double), st, same ? keep * 9 * sizeof(double) : 0));
    CU(c->f_c0.ensure(cap * 3 * sizeof(double), st, same ? keep * 3 * sizeof(double) : 0));
    CU(c->f_c1.ensure(cap * 3 * sizeof(double), st, same ? keep * 3 * sizeof(double) : 0));
    c->hit_cap = cap;
    c->hit_stride = stride;
    return JB_OK;
}

// `slice_base`: index of mol's first structure in the caller's batch (jb_query_host runs a batch
// slice by slice); hits are appended after the res->count records already in the device buffers.
static int result_alloc_block(jb_result *r, size_t cap);

static int query_run_impl(jb_context *c, const jb_templates *t, const jb_molecules *mol, const jb_query_params *q,
                          jb_result *res, int slice_base = 0, bool single_call = false)
{
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    jb_stats &S = res->stats;
    if (slice_base == 0) res->run_epoch = ++c->run_epoch;   // first slice of a run: earlier un-fetched results are overwritten
    S.pairs += (int64_t)t->T * mol->M;
    if (t->T == 0 || mol->M == 0 || t->TA == 0) return JB_OK;

    TemplateTable TT;
    TT.n_templates = t->T;
    TT.atom_off = t->d_atom_off.as<int>(); TT.spec = t->d_spec.as<int>(); TT.resseq = t->d_resseq.as<int>();
    TT.chain = t->d_chain.as<unsigned char>(); TT.pos = t->d_pos.as<double>(); TT.weight = t->d_weight.as<double>();
    TT.dist_off = t->d_dist_off.as<long long>(); TT.dist = t->d_dist.as<double>();
    TT.shell = t->d_shell.as<ShellRec>();
    TT.vol = t->d_vol.as<double>();
    SpecTable ST;
    ST.n_specs = t->S; ST.code = t->s_code.as<int>(); ST.name_off = t->s_noff.as<int>();
    ST.name_raw = t->s_nraw.as<uint32_t>(); ST.name_up = t->s_nup.as<uint32_t>();
    ST.res_off = t->s_roff.as<int>(); ST.res_up = t->s_rup.as<uint32_t>();

    const int tuple_stride = t->nmax + 1;
    // k_classify: [warps][S] counters + (when it fits) the class set and its spec masks
    // (a handful of structures: 32 warps per CTA instead of 8 -- the (class, spec) evaluations of ONE structure are the
    // critical path of a single-structure query, 66 -> ~20 us)
    int classify_threads = 256;
    if (mol->M <= 32 && ((size_t)32 * t->S + 2) * sizeof(int) + (size_t)kClassSlots * (sizeof(u64) + (size_t)((t->S + 31) / 32) * sizeof(unsigned)) <= 96 * 1024)
        classify_threads = 1024;
    const size_t cnt_bytes = ((size_t)(classify_threads / 32) * t->S + 1) / 2 * 2 * sizeof(int);
    if (cnt_bytes > 200 * 1024 || t->S > 65535)
        return fail(JB_ERR_INVALID, "%d distinct atom specs exceed the classify kernel's shared memory", t->S);
    int class_words = (t->S + 31) / 32;
    size_t classify_smem = cnt_bytes + (size_t)kClassSlots * (sizeof(u64) + (size_t)class_words * sizeof(unsigned)) +
                           (size_t)t->S * sizeof(long long) + (size_t)classify_threads * sizeof(unsigned);   // + list offsets, transposed masks
    if (classify_smem > 200 * 1024) { class_words = 0; classify_smem = cnt_bytes; }
    if (getenv("JB_CLASSIFY_PLAIN")) { class_words = 0; classify_smem = cnt_bytes; }
    if (classify_smem > 48 * 1024)
        CU(cudaFuncSetAttribute(k_classify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)classify_smem));

    CU(cudaEventRecord(c->ev0, st));
    if (!(t->shells_valid && t->shells_cutoff == q->distance_cutoff && t->shells_maxdyn == q->max_dynamic_distance)) {
        k_shells<<<t->T, 128, 0, st>>>(TT, q->distance_cutoff, q->max_dynamic_distance);
        CU(cudaGetLastError());
        S.launches++;
        t->shells_valid = true; t->shells_cutoff = q->distance_cutoff; t->shells_maxdyn = q->max_dynamic_distance;
    }

    // initial capacities (grow on demand, see the retry loop)
    if (c->fr[0].bytes == 0) {
        const size_t init = std::min<size_t>(c->ws_limit / 2, (size_t)256 << 20);
        CU(c->fr[0].ensure(init, st));
        CU(c->fr[1].ensure(init, st));
    }
    if (c->tuples.bytes == 0) CU(c->tuples.ensure((size_t)64 << 20, st));
    {
        int rc = grow_hits(c, std::max<long long>(c->hit_cap, 1 << 16), res->stride);
        if (rc != JB_OK) return rc;
    }

    // chunk the structures so that pair records and count tables stay bounded
    const long long max_pairs = 8ll << 20;
    std::vector<Range> work;
    {
        const int per = (int)std::max<long long>(1, max_pairs / std::max(1, t->T));
        for (int m0 = mol->M; m0 > 0; m0 -= per) work.push_back({ std::max(0, m0 - per), m0 });
        // processed back to front (pop_back) => ascending structure order
    }
    u64 hits_done = (u64)res->count;
    int attempts = 0;
    while (!work.empty()) {
        const Range r = work.back();
        const int Mc = r.m1 - r.m0;
        const long long pairs = (long long)Mc * t->T;
        CU(c->chead.ensure((size_t)Mc * t->S * sizeof(CandHead), st));
        CU(c->cxtab.ensure((size_t)Mc * t->S * kXBuckets * sizeof(unsigned short), st));
        CU(c->recs.ensure((size_t)pairs * sizeof(PairRec), st));
        if (c->crec.bytes == 0) {
            const long long chunk_atoms = mol->atom_off[r.m1] - mol->atom_off[r.m0];
            long long init_records = std::max<long long>(chunk_atoms * 4, 1 << 20);
            if (const char *g = getenv("JB_TEST_CREC")) init_records = std::max(64, atoi(g));   // tests: force the retry
            CU(c->crec.ensure((size_t)init_records * sizeof(float4), st));
        }
        if (c->cbin.bytes == 0) {
            // one bin per non-empty list + one per 64 further records; grown on demand like the record array
            long long init_bins = std::max<long long>((long long)(c->crec.bytes / sizeof(float4)) / kBinRecs + (long long)Mc * t->S / 4, 1024);
            if (const char *g = getenv("JB_TEST_CBIN")) init_bins = std::max(4, atoi(g));   // tests: force the retry
            CU(c->cbin.ensure((size_t)init_bins * sizeof(BinTab), st));
        }
        CU(c->cbinlist.ensure(c->cbin.bytes / sizeof(BinTab) * sizeof(int), st));
        MolTable MT;
        MT.n_molecules = Mc;
        MT.atom_off = mol->d_off.as<long long>() + r.m0;
        MT.atoms = mol->atoms.as<AtomRec>();
        MT.srec = mol->srec.as<float4>();
        MT.sname = mol->sname.as<uint32_t>(); MT.sres = mol->sres.as<uint32_t>();
        MT.ext = mol->ext.as<MolExt>() + r.m0;
        CandTable CT;
        CT.head = c->chead.as<CandHead>(); CT.rec = c->crec.as<float4>(); CT.xtab = c->cxtab.as<unsigned short>();
        CT.cap = std::min<long long>((long long)(c->crec.bytes / sizeof(float4)), (1ll << 32) - 1);   // 32-bit record offsets in the sweep
        CT.bins = c->cbin.as<BinTab>(); CT.binlist = c->cbinlist.as<int>();
        CT.bin_cap = (long long)(c->cbin.bytes / sizeof(BinTab));
        Counters *dc = c->ctr.as<Counters>();

        // reset counters; the hit cursor continues after the chunks already done
        memset(c->host_ctr, 0, sizeof(Counters));
        c->host_ctr->hits = hits_done;
        CU(cudaMemcpyAsync(dc, c->host_ctr, sizeof(Counters), cudaMemcpyHostToDevice, st));

        CU(cudaEventRecord(c->evs[0], st));
        nvtxRangePushA("jb:candidate lists (k_classify, k_xindex, k_binindex)");
        k_classify<<<Mc, classify_threads, classify_smem, st>>>(MT, ST, CT, dc, class_words);   // + the x index of its lists (K1b)
        // small chunks: ONE depth-first launch instead of one launch per level (expand_dfs.cuh); it reads the lists directly
        const bool dfs = !c->legacy && pairs <= c->dfs_pairs;
        if (c->sweep && !dfs) k_binindex<<<c->sm_count * 8, 256, 0, st>>>(CT, dc);
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->evs[1], st));
        nvtxRangePop();
        nvtxRangePushA("jb:pair set-up (k_order, k_pairs)");
        const bool pairs_small = dfs && !getenv("JB_NO_SMALL_PREPARE");
        if (pairs_small) {
            // latency path: scan order + pair record by one warp per pair (prepare_small.cuh)
            k_pairs_small<<<(unsigned)((pairs + kPairWarps - 1) / kPairWarps), kPairWarps * 32, 0, st>>>(
                Mc, TT, t->S, CT.head, MT.atom_off, q->reorder, c->recs.as<PairRec>(), dc);
            CU(cudaGetLastError());
        } else if (q->reorder == 2) {
            CU(c->ordbuf.ensure((size_t)pairs * kMaxN, st));
            k_order<<<dim3((unsigned)t->T, (unsigned)((Mc + 127) / 128)), 128, 0, st>>>(
                Mc, TT, t->S, CT.head, MT.atom_off, c->ordbuf.as<unsigned char>());
            CU(cudaGetLastError());
            S.launches++;
        }
        if (!pairs_small) {
            k_pairs<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(Mc, TT, t->S, CT.head, MT.atom_off, q->reorder,
                                                                     c->ordbuf.as<unsigned char>(), c->recs.as<PairRec>(), dc);
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(c->evs[2], st));
        nvtxRangePop();
        nvtxRangePushA("jb:frontier sweep (k_expand_grid per level) + k_superpose");

        const long long fr_ints[2] = { (long long)(c->fr[0].bytes / sizeof(int)), (long long)(c->fr[1].bytes / sizeof(int)) };
        const long long tuple_cap = (long long)(c->tuples.bytes / sizeof(int)) / tuple_stride;
        if (dfs) {
            ExpandArgs A;
            A.level = 0; A.n_specs = t->S; A.tuple_stride = tuple_stride; A.ignore_chain = q->ignore_chain;
            A.recs = c->recs.as<PairRec>(); A.shell = TT.shell; A.atoms = MT.atoms; A.ext = MT.ext;
            A.head = CT.head; A.xtab = CT.xtab; A.cand = CT.rec; A.cand_cap = CT.cap;
            A.in = nullptr; A.in_count = &dc->live_pairs; A.in_cap = pairs;
            A.out = nullptr; A.out_count = nullptr; A.out_cap = 0;
            A.tuples = c->tuples.as<int>(); A.tuple_cap = tuple_cap;
            A.ctr = dc;
            const long long want = (pairs * kRootSlices + kDfsWarps - 1) / kDfsWarps;
            k_expand_dfs_t<false><<<(unsigned)std::max<long long>(1, std::min<long long>(want, (long long)c->sm_count * 16)), kDfsThreads, 0, st>>>(A, 0, 0);
            CU(cudaGetLastError());
        }
        for (int l = 0; l < t->nmax && !dfs; l++) {
            ExpandArgs A;
            A.level = l; A.n_specs = t->S; A.tuple_stride = tuple_stride; A.ignore_chain = q->ignore_chain;
            A.recs = c->recs.as<PairRec>(); A.shell = TT.shell; A.atoms = MT.atoms; A.ext = MT.ext;
            A.head = CT.head; A.xtab = CT.xtab; A.cand = CT.rec; A.cand_cap = CT.cap;
            A.in = l == 0 ? nullptr : c->fr[(l - 1) & 1].as<int>();
            A.in_count = l == 0 ? &dc->live_pairs : &dc->level[l];
            A.in_cap = l == 0 ? pairs : fr_ints[(l - 1) & 1] / (l + 1);
            A.out = c->fr[l & 1].as<int>();
            A.out_count = &dc->level[l + 1];
            A.out_cap = fr_ints[l & 1] / (l + 2);
            A.tuples = c->tuples.as<int>(); A.tuple_cap = tuple_cap;
            A.ctr = dc;
            GridArgs GA;
            GA.bins = CT.bins; GA.bin_cap = CT.bin_cap;
            if (c->debug) {
                CU(cudaEventRecord(c->lev[l], st));
                CU(cudaMemcpyAsync(c->host_lv + 2 * l, &dc->scan_tests, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
            }
            // lanes per partial match: the configured G, widened at deep levels so that at
            // least two CTAs per SM fit their staging areas in shared memory
            int g = (l <= kMaxN && c->group_lv[l] > 0) ? c->group_lv[l] : c->group;
            // levels 0 + 1 fused: the roots are dealt out of the candidate lists by the level-1 kernel
            // itself (needs every template to have >= 2 atoms: a single-atom root is a complete tuple)
            const bool fused_root = t->nmin >= 2 && !c->legacy && !getenv("JB_NO_ROOT_FUSION");
            if (l == 0 && fused_root) continue;
            if (l == 1 && fused_root) {
                A.in = nullptr; A.in_count = &dc->live_pairs; A.in_cap = pairs;
                const int stats = c->debug != 0;
                if (c->sweep) {
                    if (g <= 1) CU((launch_expand_grid<1, 1, true>(c->sm_count, st, A, GA, stats)));
                    else if (g == 2) CU((launch_expand_grid<1, 2, true>(c->sm_count, st, A, GA, stats)));
                    else CU((launch_expand_grid<1, 4, true>(c->sm_count, st, A, GA, stats)));
                    continue;
                }
                if (g <= 1) CU((launch_expand_lv<1, 1, true>(c->sm_count, st, A, stats)));
                else if (g == 2) CU((launch_expand_lv<1, 2, true>(c->sm_count, st, A, stats)));
                else CU((launch_expand_lv<1, 4, true>(c->sm_count, st, A, stats)));
                continue;
            }
            // a small frontier is finished depth first in one launch (expand_dfs.cuh: TAIL); the level kernels below and
            // of the deeper levels then return at once
            if (l >= 3 && c->sweep && !c->legacy && c->tail_items > 0) {
                k_expand_dfs_t<true><<<(unsigned)c->sm_count * 4, kDfsThreads, 0, st>>>(A, l, c->tail_items);
                CU(cudaGetLastError());
                S.launches++;
            }
            if (l >= 1 && l <= (c->sweep ? c->grid_levels : kMaxLv) && !c->legacy) {
                if (c->sweep) CU(dispatch_expand_grid(l, g, c->sm_count, st, A, GA, c->debug != 0));
                else CU(dispatch_expand_lv(l, g, c->sm_count, st, A, c->debug != 0));
                continue;
            }
            while (g < 16 && 2 * expand_smem(g, l) > 200 * 1024) g <<= 1;
            const size_t smem = expand_smem(g, l);
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem, 1)));
            const int grid = c->sm_count * per_sm;
            const bool stats = c->debug != 0;   // scan_tests / box_survivors are diagnostics (JB_DEBUG)
            cudaError_t e;
            switch (g) {
            case 1: e = stats ? launch_expand<1, true>(grid, smem, st, A) : launch_expand<1, false>(grid, smem, st, A); break;
            case 2: e = stats ? launch_expand<2, true>(grid, smem, st, A) : launch_expand<2, false>(grid, smem, st, A); break;
            case 4: e = stats ? launch_expand<4, true>(grid, smem, st, A) : launch_expand<4, false>(grid, smem, st, A); break;
            case 8: e = stats ? launch_expand<8, true>(grid, smem, st, A) : launch_expand<8, false>(grid, smem, st, A); break;
            default: e = stats ? launch_expand<16, true>(grid, smem, st, A) : launch_expand<16, false>(grid, smem, st, A); break;
            }
            CU(e);
        }
        CU(cudaEventRecord(c->evs[3], st));
        if (c->debug) {
            CU(cudaEventRecord(c->lev[t->nmax], st));
            CU(cudaMemcpyAsync(c->host_lv + 2 * t->nmax, &dc->scan_tests, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        }
        HitBuf HB;
        HB.pair = q->hit_order != JB_ORDER_CANONICAL ? c->h_pair.as<int>() : nullptr;
        HB.mol = c->h_mol.as<int>(); HB.tpl = c->h_tpl.as<int>(); HB.atoms = c->h_atoms.as<int>();
        HB.rmsd = c->h_rmsd.as<double>(); HB.rot = c->h_rot.as<double>(); HB.c0 = c->h_c0.as<double>();
        HB.c1 = c->h_c1.as<double>(); HB.cap = c->hit_cap; HB.stride = c->hit_stride;
        k_superpose<<<c->sm_count * 8, 128, 0, st>>>(c->tuples.as<int>(), &dc->tuple_slots, tuple_cap, tuple_stride,
                                                    c->recs.as<PairRec>(), TT, MT, slice_base + r.m0, q->rmsd_threshold, q->all_tuples, HB,
                                                    dc);
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->evs[4], st));
        CU(cudaMemcpyAsync(c->host_ctr, dc, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        nvtxRangePop();
        const Counters &hc = *c->host_ctr;
        attempts++;
        if (c->debug) {
            fprintf(stderr, "[jb] chunk [%d,%d) live=%llu cand=%llu tests=%llu scanned=%llu box=%llu tuples=%llu\n", r.m0,
                    r.m1, hc.live_pairs, hc.cand_total, hc.cand_tests, hc.scan_tests, hc.box_pass, hc.tuples);
            for (int l = 0; l < t->nmax; l++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, c->lev[l], c->lev[l + 1]);
                fprintf(stderr, "[jb]   level %2d: in=%12llu out=%12llu  %8.3f ms  covered=%12llu box=%12llu\n", l,
                        l == 0 ? hc.live_pairs : hc.level[l], hc.level[l + 1], ms, c->host_lv[2 * l + 2] - c->host_lv[2 * l],
                        c->host_lv[2 * l + 3] - c->host_lv[2 * l + 1]);
            }
        }
        S.launches += dfs ? 3 : 3 + t->nmax - ((t->nmin >= 2 && !c->legacy && !getenv("JB_NO_ROOT_FUSION")) ? 1 : 0);
        {
            float ms[4] = { 0, 0, 0, 0 };
            for (int i = 0; i < 4; i++) CU(cudaEventElapsedTime(&ms[i], c->evs[i], c->evs[i + 1]));
            S.classify_ms += ms[0]; S.pairs_ms += ms[1]; S.expand_ms += ms[2]; S.superpose_ms += ms[3];
        }

        // what this chunk really used of the grow-only buffers (context_decay releases the ones left idle)
        {
            auto used = [](DevBuf &b, size_t bytes) { b.recent = std::max(b.recent, std::min(bytes, b.bytes)); };
            used(c->crec, (size_t)hc.cand_total * sizeof(float4));
            used(c->cbin, (size_t)hc.bin_total * sizeof(BinTab));
            used(c->cbinlist, (size_t)hc.bin_total * sizeof(int));
            // (levels from tail_level on were finished depth first: their counters are work counters, no frontier was written)
            const int fr_levels = hc.tail_level ? std::min<int>(t->nmax, (int)hc.tail_level) : t->nmax;
            for (int l = 0; l < fr_levels; l++) used(c->fr[l & 1], (size_t)hc.level[l + 1] * (l + 2) * sizeof(int));
            used(c->tuples, (size_t)hc.tuple_slots * tuple_stride * sizeof(int));
            const size_t nhits = (size_t)std::min<u64>(hc.hits, (u64)c->hit_cap);   // cumulative over the run
            DevBuf *hb[] = { &c->h_mol, &c->h_tpl, &c->h_atoms, &c->h_rmsd, &c->h_rot, &c->h_c0, &c->h_c1, &c->f_mol, &c->f_tpl,
                             &c->f_atoms, &c->f_rmsd, &c->f_rot, &c->f_c0, &c->f_c1, &c->h_pair, &c->h_keys };
            for (DevBuf *b : hb)
                if (c->hit_cap > 0) used(*b, (size_t)((double)b->bytes * (double)nhits / (double)c->hit_cap));
        }

        // ---- capacity checks: grow and retry, or split the chunk --------------------
        bool retry = false, split = false;
        if ((long long)hc.cand_total > CT.cap) {
            const size_t need = (size_t)hc.cand_total + (size_t)hc.cand_total / 8;
            CU(c->crec.ensure(need * sizeof(float4), st));
            retry = true;
        }
        if (c->sweep && (long long)hc.bin_total > CT.bin_cap) {
            const size_t need = (size_t)hc.bin_total + (size_t)hc.bin_total / 8;
            CU(c->cbin.ensure(need * sizeof(BinTab), st));
            retry = true;
        }
        if (!retry) {
            long long need_ints[2] = { 0, 0 };
            const int fr_levels = hc.tail_level ? std::min<int>(t->nmax, (int)hc.tail_level) : t->nmax;
            for (int l = 0; l < fr_levels && !dfs; l++)      // (the depth-first sweeps keep no frontier)
                need_ints[l & 1] = std::max(need_ints[l & 1], (long long)hc.level[l + 1] * (l + 2));
            for (int b = 0; b < 2; b++)
                if (need_ints[b] > fr_ints[b]) {
                    const size_t want = (size_t)(need_ints[b] + need_ints[b] / 2) * sizeof(int);
                    if (want * 2 > c->ws_limit && Mc > 1) split = true;
                    else if (want * 2 > c->ws_limit)
                        return fail(JB_ERR_CAPACITY, "structure %d needs a %zu MiB frontier; raise workspace_bytes", r.m0, want >> 20);
                    else { CU(c->fr[b].ensure(want, st)); retry = true; }
                }
            if ((long long)hc.tuple_slots > tuple_cap) {     // (records reserved: the tuples plus the holes of the last chunks)
                CU(c->tuples.ensure((size_t)(hc.tuple_slots + hc.tuple_slots / 2) * tuple_stride * sizeof(int), st));
                retry = true;
            }
            if ((long long)hc.hits > c->hit_cap) {
                int rc = grow_hits(c, (long long)hc.hits, res->stride);
                if (rc != JB_OK) return rc;
                retry = true;
            }
        }
        if (split) {
            work.pop_back();
            const int mid = r.m0 + Mc / 2;
            work.push_back({ mid, r.m1 });
            work.push_back({ r.m0, mid });
            S.retries++;
            continue;
        }
        if (retry) { S.retries++; continue; }

        // ---- chunk done: put its hits into canonical order on the device ------------------
        work.pop_back();
        NvtxRange order_range("jb:hit ordering + gather");
        bool ordered = false;
        if (hc.hits > hits_done && (long long)(hc.hits - hits_done) <= kSmallHits && !getenv("JB_ORDER_GENERAL")) {
            // ---- small result: ONE kernel orders the records and writes them out -- straight into the page-locked
            // result block when this chunk is the whole run (no device -> host copies left for jb_result_fetch)
            const long long h0 = (long long)hits_done, h1 = (long long)hc.hits, nh = h1 - h0;
            const bool to_host = single_call && h0 == 0 && r.m0 == 0 && r.m1 == mol->M;
            HitBuf DB = HB;
            DB.mol = c->f_mol.as<int>(); DB.tpl = c->f_tpl.as<int>(); DB.atoms = c->f_atoms.as<int>();
            DB.rmsd = c->f_rmsd.as<double>(); DB.rot = c->f_rot.as<double>(); DB.c0 = c->f_c0.as<double>();
            DB.c1 = c->f_c1.as<double>();
            double *xyz = nullptr;
            if (to_host) {
                int rc = result_alloc_block(res, (size_t)nh);
                if (rc != JB_OK) return rc;
                DB.mol = res->mol; DB.tpl = res->tpl; DB.atoms = res->atoms; DB.rmsd = res->rmsd; DB.rot = res->rot;
                DB.c0 = res->c0; DB.c1 = res->c1;
                xyz = res->xyz;
            } else if (res->coords) {
                CU(c->f_xyz.ensure((size_t)c->hit_cap * c->hit_stride * 3 * sizeof(double), st, (size_t)h0 * c->hit_stride * 3 * sizeof(double)));
                xyz = c->f_xyz.as<double>();
            }
            int *flag = reinterpret_cast<int *>(&c->o_work.as<RunWork>()[1].max_npad);
            const bool ranked = q->hit_order != JB_ORDER_CANONICAL;
            if (ranked) CU(cudaMemsetAsync(flag, 0, sizeof(int), st));
            k_hits_small<<<1, 256, 0, st>>>(h0, (int)nh, c->hit_stride, HB, DB, nullptr, flag);
            if (res->coords)
                k_hit_coords<<<(unsigned)((nh * 32 + 255) / 256), 256, 0, st>>>(h0, h1, c->hit_stride, slice_base + r.m0, MT, DB, xyz);
            CU(cudaGetLastError());
            S.launches += res->coords ? 2 : 1;
            int multi = 0;
            if (ranked) {                                   // with one hit per pair at most the order inside a pair is moot
                CU(cudaMemcpyAsync(&multi, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
            }
            ordered = multi == 0;
            if (ordered && to_host) res->direct = true;
        }
        if (hc.hits > hits_done && !ordered) {
            const long long h0 = (long long)hits_done, h1 = (long long)hc.hits, nh = h1 - h0;
            const int n_tiles = (int)((pairs + 1 + kScanTile - 1) / kScanTile);
            CU(c->o_off.ensure((size_t)(pairs + 1) * sizeof(unsigned), st));
            CU(c->o_fill.ensure((size_t)pairs * sizeof(unsigned), st));
            CU(c->o_grouped.ensure((size_t)nh * sizeof(unsigned), st));
            CU(c->o_order.ensure((size_t)nh * sizeof(unsigned), st));
            CU(c->o_tiles.ensure((size_t)n_tiles * sizeof(unsigned), st));
            CU(cudaMemsetAsync(c->o_off.p, 0, (size_t)(pairs + 1) * sizeof(unsigned), st));
            CU(cudaMemsetAsync(c->o_fill.p, 0, (size_t)pairs * sizeof(unsigned), st));
            HitOrderArgs OA;
            OA.h0 = h0; OA.h1 = h1; OA.mol_base = slice_base + r.m0; OA.n_templates = t->T; OA.stride = c->hit_stride;
            OA.mol = HB.mol; OA.tpl = HB.tpl; OA.atoms = HB.atoms;
            bool ranked = q->hit_order != JB_ORDER_CANONICAL;
            if (ranked && nh <= 1024) {
                // sparse small result: with at most one hit per pair the order inside a pair is moot
                int *flag = reinterpret_cast<int *>(&c->o_work.as<RunWork>()[1].max_npad);
                CU(cudaMemsetAsync(flag, 0, sizeof(int), st));
                k_hit_multi<<<1, 256, 0, st>>>(h0, (int)nh, HB.mol, HB.tpl, flag);
                int multi = 1;
                CU(cudaMemcpyAsync(&multi, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                S.launches++;
                ranked = multi != 0;
            }
            if (ranked) {
                // rank vectors of the reference's depth-first order replace the atom tuples as sort keys
                int rc = build_reference_keys(c, t, mol, q, r.m0, Mc, MT, CT, TT, h0, h1, S);
                if (rc != JB_OK) return rc;
                OA.atoms = c->h_keys.as<int>();
            }
            OA.off = c->o_off.as<unsigned>(); OA.fill = c->o_fill.as<unsigned>();
            OA.grouped = c->o_grouped.as<unsigned>(); OA.order = c->o_order.as<unsigned>();
            OA.runid = nullptr;
            const unsigned hb = (unsigned)((nh + 255) / 256);
            // keyed path: one radix sort groups the hits by pair AND orders them by the first two key
            // columns; the run machinery below then only resolves ties.  Needs the columns to fit
            // their 20/21 bits (atom indices / ranks < 2^20) and the pair its 23 bits.
            long long amax = 0;
            for (int m = r.m0; m < r.m1; m++) amax = std::max(amax, mol->atom_off[m + 1] - mol->atom_off[m]);
            const bool keyed = amax < (1ll << kKeyCol0Bits) && pairs <= (1ll << 23) && !getenv("JB_ORDER_COUNTING");
            long long n_runs_bound = pairs;
            const unsigned *d_n_runs = nullptr;
            if (keyed) {
                CU(c->k_key.ensure((size_t)nh * sizeof(u64), st));   CU(c->k_key2.ensure((size_t)nh * sizeof(u64), st));
                CU(c->k_iota.ensure((size_t)nh * sizeof(unsigned), st));
                CU(c->k_runid.ensure((size_t)(nh + 1) * sizeof(unsigned), st));
                CU(c->o_off.ensure((size_t)(nh + 2) * sizeof(unsigned), st));
                OA.off = c->o_off.as<unsigned>();
                k_hit_key<<<hb, 256, 0, st>>>(OA, c->k_key.as<u64>(), c->k_iota.as<unsigned>());
                int pair_bits = 1;
                while ((1ll << pair_bits) < pairs) pair_bits++;
                size_t tmp = 0;
                const int end_bit = kKeyCol0Bits + kKeyCol1Bits + pair_bits;
                CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->k_key.as<u64>(), c->k_key2.as<u64>(), c->k_iota.as<unsigned>(),
                                                   OA.grouped, (int)nh, 0, end_bit, st));
                CU(c->q_cub.ensure(tmp, st));
                CU(cub::DeviceRadixSort::SortPairs(c->q_cub.p, tmp, c->k_key.as<u64>(), c->k_key2.as<u64>(),
                                                   c->k_iota.as<unsigned>(), OA.grouped, (int)nh, 0, end_bit, st));
                // runs of equal keys: flags -> exclusive scan -> run ids / offsets / count
                unsigned *flag = c->k_runid.as<unsigned>();           // scanned in place, then overwritten by the ids
                const int ft = (int)((nh + 1 + kScanTile - 1) / kScanTile);
                CU(c->o_tiles.ensure((size_t)std::max(ft, n_tiles) * sizeof(unsigned), st));
                CU(c->k_scan.ensure((size_t)(nh + 1) * sizeof(unsigned), st));
                unsigned *scanned = c->k_scan.as<unsigned>();
                k_run_flags<<<(unsigned)((nh + 1 + 255) / 256), 256, 0, st>>>(nh, c->k_key2.as<u64>(), scanned);
                k_scan_tiles<<<ft, kScanThreads, 0, st>>>(scanned, nh + 1, c->o_tiles.as<unsigned>());
                k_scan_totals<<<1, kScanThreads, 0, st>>>(c->o_tiles.as<unsigned>(), ft);
                k_scan_add<<<ft, kScanThreads, 0, st>>>(scanned, nh + 1, c->o_tiles.as<unsigned>());
                k_run_heads<<<hb, 256, 0, st>>>(nh, c->k_key2.as<u64>(), scanned, flag, OA.off, &c->o_work.as<RunWork>()[1].n_runs);
                OA.runid = flag;
                d_n_runs = &c->o_work.as<RunWork>()[1].n_runs;
                n_runs_bound = nh;
                k_hit_rank<<<hb, 256, 0, st>>>(OA);
                S.launches += 7;   // + the radix sort passes (library kernels)
            } else {
                k_hit_hist<<<hb, 256, 0, st>>>(OA);
                k_scan_tiles<<<n_tiles, kScanThreads, 0, st>>>(OA.off, pairs + 1, c->o_tiles.as<unsigned>());
                k_scan_totals<<<1, kScanThreads, 0, st>>>(c->o_tiles.as<unsigned>(), n_tiles);
                k_scan_add<<<n_tiles, kScanThreads, 0, st>>>(OA.off, pairs + 1, c->o_tiles.as<unsigned>());
                k_hit_group<<<hb, 256, 0, st>>>(OA);
                k_hit_rank<<<hb, 256, 0, st>>>(OA);
                S.launches += 7;
            }
            if (nh > (long long)kShortRun) {
                // dense-hit pairs: segmented tiled bitonic sort of the long runs (hits.cuh)
                const unsigned tile_cap = (unsigned)(nh / kTile + std::min<long long>(n_runs_bound, nh / kShortRun + 1) + 1);
                CU(c->o_runs.ensure((size_t)tile_cap * sizeof(RunTile), st));
                CU(cudaMemsetAsync(c->o_work.p, 0, sizeof(RunWork), st));
                k_run_tiles<<<(unsigned)((n_runs_bound + 255) / 256), 256, 0, st>>>(OA, n_runs_bound, d_n_runs, c->o_runs.as<RunTile>(),
                                                                                    tile_cap, c->o_work.as<RunWork>());
                CU(cudaMemcpyAsync(c->host_work, c->o_work.p, sizeof(RunWork), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                const unsigned nt = c->host_work->n_tiles, max_npad = c->host_work->max_npad;
                if (nt > tile_cap) return fail(JB_ERR_CAPACITY, "hit ordering: %u tiles exceed the work list (%u)", nt, tile_cap);
                S.launches++;
                const bool force_lsd = getenv("JB_ORDER_LSD") != nullptr;       // tests: take the dense path at any size
                // ... worth it when a good share of the hits sits in long runs (n log^2 n per run otherwise)
                const bool dense = nt > 0 && (long long)c->host_work->long_hits * 8 >= nh;
                if ((dense || force_lsd) && keyed && !getenv("JB_ORDER_NO_LSD")) {
                    // long tie runs: order the whole chunk by a multi-pass LSD radix sort (hits.cuh)
                    int bits = 1;
                    while ((1ll << bits) < amax + 1) bits++;                  // columns < amax (or all ones: padding)
                    bits = std::min(bits + 1, (int)kKeyCol1Bits);
                    const int per = 64 / bits;
                    unsigned *perm = c->k_iota.as<unsigned>(), *perm2 = OA.grouped;
                    k_iota<<<hb, 256, 0, st>>>(perm, nh);
                    size_t tmp = 0;
                    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->k_key.as<u64>(), c->k_key2.as<u64>(), perm, perm2,
                                                       (int)nh, 0, 64, st));
                    CU(c->q_cub.ensure(tmp, st));
                    for (int hi = c->hit_stride; hi > 2;) {
                        const int lo = std::max(2, hi - per);
                        k_lsd_key<<<hb, 256, 0, st>>>(OA, perm, lo, hi, bits, 0, c->k_key.as<u64>());
                        CU(cub::DeviceRadixSort::SortPairs(c->q_cub.p, tmp, c->k_key.as<u64>(), c->k_key2.as<u64>(), perm, perm2,
                                                           (int)nh, 0, (hi - lo) * bits, st));
                        std::swap(perm, perm2);
                        S.launches += 1;
                        hi = lo;
                    }
                    int pair_bits = 1;
                    while ((1ll << pair_bits) < pairs) pair_bits++;
                    k_lsd_key<<<hb, 256, 0, st>>>(OA, perm, 0, 0, bits, 1, c->k_key.as<u64>());
                    CU(cub::DeviceRadixSort::SortPairs(c->q_cub.p, tmp, c->k_key.as<u64>(), c->k_key2.as<u64>(), perm, OA.order,
                                                       (int)nh, 0, kKeyCol0Bits + kKeyCol1Bits + pair_bits, st));
                    S.launches += 2;
                } else if (nt > 0) {
                    const RunTile *rt = c->o_runs.as<RunTile>();
                    k_tile_sort<<<nt, kTileThreads, 0, st>>>(OA, rt);
                    S.launches++;
                    for (unsigned k = 2u * kTile; k <= max_npad && k != 0; k <<= 1) {
                        for (unsigned j = k >> 1; j >= (unsigned)kTile; j >>= 1) {
                            k_run_stage<<<nt, kTileThreads, 0, st>>>(OA, rt, k, j);
                            S.launches++;
                        }
                        k_tile_merge<<<nt, kTileThreads, 0, st>>>(OA, rt, k);
                        S.launches++;
                    }
                    k_run_publish<<<nt, kTileThreads, 0, st>>>(OA, rt);
                    S.launches++;
                }
            }
            HitBuf FB = HB;
            FB.mol = c->f_mol.as<int>(); FB.tpl = c->f_tpl.as<int>(); FB.atoms = c->f_atoms.as<int>();
            FB.rmsd = c->f_rmsd.as<double>(); FB.rot = c->f_rot.as<double>(); FB.c0 = c->f_c0.as<double>();
            FB.c1 = c->f_c1.as<double>();
            k_hit_gather<<<(unsigned)((nh * 32 + 255) / 256), 256, 0, st>>>(h0, h1, c->hit_stride, OA.order, HB, FB);
            CU(cudaGetLastError());
            if (res->coords) {
                // Hit.atoms(transform=True) for every ordered hit of the chunk, while its structures are resident
                CU(c->f_xyz.ensure((size_t)c->hit_cap * c->hit_stride * 3 * sizeof(double), st,
                                   (size_t)h0 * c->hit_stride * 3 * sizeof(double)));
                k_hit_coords<<<(unsigned)((nh * 32 + 255) / 256), 256, 0, st>>>(h0, h1, c->hit_stride, slice_base + r.m0, MT, FB,
                                                                              c->f_xyz.as<double>());
                CU(cudaGetLastError());
                S.launches++;
            }
        }
        hits_done = hc.hits;
        S.chunks++;
        S.live_pairs += (int64_t)hc.live_pairs;
        S.candidate_atoms += (int64_t)hc.cand_atoms;
        S.candidate_records += (int64_t)hc.cand_total;
        S.candidate_tests += (int64_t)hc.cand_tests;
        S.scan_tests += (int64_t)hc.scan_tests;
        S.box_survivors += (int64_t)hc.box_pass;
        S.tuples += (int64_t)hc.tuples;
        S.nan_dropped += (int64_t)hc.nan_dropped;
        S.frontier_items += (int64_t)hc.tuples;
        S.frontier_ints += (int64_t)hc.tuple_ints;
        for (int l = 1; l <= t->nmax; l++) {
            S.frontier_items += (int64_t)hc.level[l];
            S.frontier_ints += (int64_t)hc.level[l] * l;
        }
    }
    CU(cudaEventRecord(c->ev1, st));
    CU(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    S.kernel_ms += ms;
    S.hits = (int64_t)hits_done;
    res->count = (long long)hits_done;
    (void)attempts;
    return JB_OK;
}

static int query_run_once(jb_context *c, const jb_templates *t, const jb_molecules *mol, const jb_query_params *q,
                          jb_result **out);

// Out of device memory: the grow-only buffers of the context may still be sized for an earlier, larger
// workload (or hit volume) -- give them back and try once more before reporting JB_ERR_NOMEM.
extern "C" int jb_query_run(jb_context *c, const jb_templates *t, const jb_molecules *mol, const jb_query_params *q,
                            jb_result **out)
{
    if (c) context_decay(c);
    int rc = query_run_once(c, t, mol, q, out);
    if (rc == JB_ERR_NOMEM && c) {
        context_trim(c);
        rc = query_run_once(c, t, mol, q, out);
    }
    return rc;
}

static int query_run_once(jb_context *c, const jb_templates *t, const jb_molecules *mol, const jb_query_params *q,
                          jb_result **out)
{
    if (!c || !t || !mol || !q || !out) return fail(JB_ERR_INVALID, "jb_query_run: NULL argument");
    *out = nullptr;
    if (t->ctx != c || mol->ctx != c) return fail(JB_ERR_INVALID, "templates/molecules belong to another context");
    if (q->ignore_chain < 0 || q->ignore_chain > 2) return fail(JB_ERR_INVALID, "invalid ignore_chain %d", q->ignore_chain);
    if (q->reorder < 0 || q->reorder > 2) return fail(JB_ERR_INVALID, "invalid reorder %d", q->reorder);
    if (q->hit_order < 0 || q->hit_order > 2) return fail(JB_ERR_INVALID, "invalid hit_order %d", q->hit_order);
    jb_result *res = new jb_result();
    res->ctx = c;
    res->stride = std::max(t->nmax, 1);
    res->best_match = q->all_tuples ? 0 : q->best_match;
    res->n_templates = t->T;
    res->trim_epoch = c->trim_epoch;
    res->coords = (q->flags & JB_QUERY_COORDS) ? 1 : 0;
    const int rc = query_run_impl(c, t, mol, q, res, 0, true);
    if (rc != JB_OK) { delete res; return rc; }
    *out = res;
    return JB_OK;
}

static size_t align64(size_t v) { return (v + 63) & ~(size_t)63; }

// A page-locked result block goes back to the context's pool (caller holds pool_mu).  The pool keeps the
// kPinPool LARGEST blocks: pinning memory costs ~0.2 ms per MiB, so a workload with large hit volumes must
// find its blocks again even after many small queries went through the same context (a pool that refused
// newcomers when full made every cfg 4 call page-lock and release a fresh 1.3 GB block: 230 -> 450 ms).
// Blocks of at most kPinSmall bytes (single-structure queries: every block is allocated at least that large, so
// they are interchangeable) are pooled apart, so that neither kind evicts the other.
constexpr size_t kPinPool = 8, kPinSmall = (size_t)1 << 16;
static void pool_put(jb_context *c, void *block, size_t bytes)
{
    size_t n_small = 0, n_large = 0;
    for (auto &b : c->pin_pool) (b.second <= kPinSmall ? n_small : n_large)++;
    if (bytes <= kPinSmall) {
        if (n_small < kPinPool) c->pin_pool.emplace_back(block, bytes);
        else cudaFreeHost(block);
        return;
    }
    c->pin_pool.emplace_back(block, bytes);
    if (n_large + 1 > kPinPool) {
        size_t small = c->pin_pool.size();
        for (size_t i = 0; i < c->pin_pool.size(); i++)
            if (c->pin_pool[i].second > kPinSmall && (small == c->pin_pool.size() || c->pin_pool[i].second < c->pin_pool[small].second))
                small = i;
        cudaFreeHost(c->pin_pool[small].first);
        c->pin_pool.erase(c->pin_pool.begin() + small);
    }
}

// One page-locked block for the seven hit arrays, laid out for `cap` records and recycled
// through the context's pool.
static int result_alloc_block(jb_result *r, size_t cap)
{
    jb_context *c = r->ctx;
    const size_t stride = (size_t)r->stride;
    const size_t o_mol = 0, o_tpl = o_mol + align64(cap * 4), o_atoms = o_tpl + align64(cap * 4),
                 o_rmsd = o_atoms + align64(cap * stride * 4), o_rot = o_rmsd + align64(cap * 8),
                 o_c0 = o_rot + align64(cap * 72), o_c1 = o_c0 + align64(cap * 24), o_xyz = o_c1 + align64(cap * 24),
                 total = o_xyz + (r->coords ? align64(cap * stride * 24) : 0) + 64;
    {
        std::lock_guard<std::mutex> g(c->pool_mu);
        if (r->block) {   // a smaller block from the pipelined path: back to the pool
            pool_put(c, r->block, r->block_bytes);
            r->block = nullptr;
        }
        size_t best = c->pin_pool.size();                // smallest block that fits
        for (size_t i = 0; i < c->pin_pool.size(); i++)
            if (c->pin_pool[i].second >= total && (best == c->pin_pool.size() || c->pin_pool[i].second < c->pin_pool[best].second))
                best = i;
        if (best < c->pin_pool.size()) {
            r->block = c->pin_pool[best].first; r->block_bytes = c->pin_pool[best].second;
            c->pin_pool.erase(c->pin_pool.begin() + best);
        }
    }
    if (!r->block) {
        const size_t want = std::max(total + total / 4, kPinSmall);
        CU(cudaMallocHost(&r->block, want));
        r->block_bytes = want;
    }
    char *base = static_cast<char *>(r->block);
    r->mol = reinterpret_cast<int32_t *>(base + o_mol); r->tpl = reinterpret_cast<int32_t *>(base + o_tpl);
    r->atoms = reinterpret_cast<int32_t *>(base + o_atoms); r->rmsd = reinterpret_cast<double *>(base + o_rmsd);
    r->rot = reinterpret_cast<double *>(base + o_rot); r->c0 = reinterpret_cast<double *>(base + o_c0);
    r->c1 = reinterpret_cast<double *>(base + o_c1);
    r->xyz = r->coords ? reinterpret_cast<double *>(base + o_xyz) : nullptr;
    r->block_cap = cap;
    return JB_OK;
}

// device -> host copy of the ordered hit records [from, to) on stream `st` (no sync)
static int result_copy_range(jb_result *r, size_t from, size_t to, cudaStream_t st)
{
    jb_context *c = r->ctx;
    if (to <= from) return JB_OK;
    const size_t n = to - from, stride = (size_t)r->stride;
#define D2H(dst, buf, T, w) CU(cudaMemcpyAsync((dst) + from * (w), (buf).template as<T>() + from * (w), n * (w) * sizeof(T), cudaMemcpyDeviceToHost, st))
    D2H(r->mol, c->f_mol, int, 1);
    D2H(r->tpl, c->f_tpl, int, 1);
    D2H(r->atoms, c->f_atoms, int, stride);
    D2H(r->rmsd, c->f_rmsd, double, 1);
    D2H(r->rot, c->f_rot, double, 9);
    D2H(r->c0, c->f_c0, double, 3);
    D2H(r->c1, c->f_c1, double, 3);
    if (r->coords) D2H(r->xyz, c->f_xyz, double, stride * 3);
#undef D2H
    return JB_OK;
}

// keep the smallest RMSD per (molecule, template), compacting in place; ties -> first in the
// record order (the reference keeps the first in ITS iteration order, _jess.pyx:2238)
static void result_finalize(jb_result *r)
{
    const size_t n = (size_t)r->count, stride = (size_t)r->stride;
    r->kept = (long long)n;
    if (r->best_match && n) {
        size_t w = 0;
        for (size_t i = 0; i < n;) {
            size_t j = i, best = i;
            while (j < n && r->mol[j] == r->mol[i] && r->tpl[j] == r->tpl[i]) {
                if (r->rmsd[j] < r->rmsd[best]) best = j;
                j++;
            }
            r->mol[w] = r->mol[best]; r->tpl[w] = r->tpl[best]; r->rmsd[w] = r->rmsd[best];
            memmove(&r->atoms[w * stride], &r->atoms[best * stride], stride * sizeof(int32_t));
            memmove(&r->rot[w * 9], &r->rot[best * 9], 9 * sizeof(double));
            memmove(&r->c0[w * 3], &r->c0[best * 3], 3 * sizeof(double));
            memmove(&r->c1[w * 3], &r->c1[best * 3], 3 * sizeof(double));
            if (r->coords) memmove(&r->xyz[w * stride * 3], &r->xyz[best * stride * 3], stride * 3 * sizeof(double));
            w++;
            i = j;
        }
        r->kept = (long long)w;
    }
    r->fetched = true;
}

extern "C" int jb_result_fetch(jb_result *r, jb_hits *out)
{
    if (!r || !out) return fail(JB_ERR_INVALID, "jb_result_fetch: NULL argument");
    jb_context *c = r->ctx;
    CU(cudaSetDevice(c->device));
    if (!r->fetched && r->direct) {                       // the records are in the host block already (k_hits_small)
        CU(cudaStreamSynchronize(c->stream));
        result_finalize(r);
    }
    if (!r->fetched) {
        if (r->trim_epoch != c->trim_epoch)
            return fail(JB_ERR_INVALID, "jb_result_fetch: the hits of this run were discarded by jb_context_trim");
        if (r->run_epoch != c->run_epoch)
            return fail(JB_ERR_INVALID, "jb_result_fetch: a later run on this context has overwritten the hits of this one "
                                        "(fetch a result before the next jb_query_run / jb_query_host)");
        const double t_0 = now_ms();
        const size_t n = (size_t)r->count;
        int rc = result_alloc_block(r, n);
        if (rc != JB_OK) return rc;
        rc = result_copy_range(r, 0, n, c->stream);
        if (rc != JB_OK) return rc;
        CU(cudaStreamSynchronize(c->stream));
        result_finalize(r);
        if (c->debug) fprintf(stderr, "[jb] fetch %zu hits: %.2f ms\n", n, now_ms() - t_0);
    }
    out->count = (int64_t)r->kept;
    out->stride = r->stride;
    out->reserved = 0;
    out->molecule = r->mol; out->template_index = r->tpl; out->atoms = r->atoms;
    out->rmsd = r->rmsd; out->rotation = r->rot; out->centre0 = r->c0; out->centre1 = r->c1;
    out->coords = r->coords ? r->xyz : nullptr;
    return JB_OK;
}

extern "C" int jb_result_stats(jb_result *r, jb_stats *out)
{
    if (!r || !out) return fail(JB_ERR_INVALID, "jb_result_stats: NULL argument");
    *out = r->stats;
    return JB_OK;
}

extern "C" void jb_result_free(jb_result *r)
{
    if (!r) return;
    if (r->block) {   // recycled by the next fetch (bounded pool); may run while another thread queries
        std::lock_guard<std::mutex> g(r->ctx->pool_mu);
        pool_put(r->ctx, r->block, r->block_bytes);
    }
    delete r;
}

// jb_query_host for large batches: the structures are processed in slices so that the host ->
// device copy of slice k+1 (copy stream, double-buffered staging) and the device -> host copy of
// slice k-1's hits overlap the kernels of slice k (compute stream).  Record order is unchanged:
// slices are ascending runs of structures.
static int query_host_pipelined(jb_context *c, const jb_templates *t, const HostSource &src,
                                const jb_query_params *q, int n_slices, jb_result **out, jb_hits *hits)
{
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; i++)
        if (!c->slice_mol[i]) c->slice_mol[i] = new jb_molecules();
    // slice boundaries: equal shares of the atoms
    const int M = src.M;
    const int64_t *aoff = src.offsets();
    const long long NA = aoff[M];
    std::vector<int> cut{ 0 };
    for (int k = 1; k < n_slices; k++) {
        // records mode: the packing of the FIRST slice is the only one nothing hides -- keep it half as large
        // two slices of a packed batch: the first takes 20 % of the atoms (JB_FIRST_SLICE) -- its upload is the one nothing
        // hides; measured on cfg 3 end to end: 50 % 10.16 ms, 40 % 10.08, 30 % 10.02, 20 % 9.81, 15 % 9.89, 10 % 9.94 (r3z, r4a)
        const int64_t want = !src.soa ? NA * (2 * k - 1) / (2 * n_slices - 1)
                             : (n_slices == 2 && c->first_slice_pct > 0) ? NA * c->first_slice_pct / 100 : NA * k / n_slices;
        int m = (int)(std::lower_bound(aoff, aoff + M + 1, want) - aoff);
        m = std::min(std::max(m, cut.back() + 1), M - (n_slices - k));
        cut.push_back(m);
    }
    cut.push_back(M);
    jb_result *res = new jb_result();
    res->ctx = c;
    res->stride = std::max(t->nmax, 1);
    res->best_match = q->all_tuples ? 0 : q->best_match;
    res->n_templates = t->T;
    res->trim_epoch = c->trim_epoch;
    res->coords = (q->flags & JB_QUERY_COORDS) ? 1 : 0;
    int rc = result_alloc_block(res, (size_t)std::max<long long>(c->last_hits + c->last_hits / 8, 1 << 16));
    bool direct = rc == JB_OK;                        // hits still fit the block laid out above
    size_t copied = 0;
    // Slice k + 1 is staged (records mode: packed on host threads, then copied) by a helper thread while this thread
    // drives the kernels of slice k.  Both the device image slice_mol[(k + 1) & 1] and the host block host_stage[(k + 1) & 1]
    // were last used by slice k - 1, whose upload and kernels are complete when iteration k starts.
    const double t_begin = now_ms();
    auto stage = [&](int k) -> std::pair<int, std::string> {
        cudaSetDevice(c->device);
        const double t_0 = now_ms();
        int r = stage_source(c, c->slice_mol[k & 1], src, cut[k], cut[k + 1], c->copy_stream, c->host_stage[k & 1]);
        if (c->debug >= 2) fprintf(stderr, "[jb]   slice %d staged (structures [%d, %d)): %.2f - %.2f ms\n", k, cut[k], cut[k + 1],
                                   t_0 - t_begin, now_ms() - t_begin);
        if (r == JB_OK && cudaEventRecord(c->ev_up[k & 1], c->copy_stream) != cudaSuccess) r = fail(JB_ERR_CUDA, "cudaEventRecord failed");
        return { r, r == JB_OK ? std::string() : std::string(g_err) };
    };
    if (rc == JB_OK) rc = stage(0).first;
    for (int k = 0; k < n_slices && rc == JB_OK; k++) {
        jb_molecules *cur = c->slice_mol[k & 1];
        std::future<std::pair<int, std::string>> next;
        if (k + 1 < n_slices) {
            if (src.soa) {                               // copies only: enqueue them right here
                rc = stage(k + 1).first;
                if (rc != JB_OK) break;
            } else next = std::async(std::launch::async, stage, k + 1);
        }
        int rk = JB_OK;
        const double t_run = now_ms();
        if (cudaStreamWaitEvent(c->stream, c->ev_up[k & 1], 0) != cudaSuccess) rk = fail(JB_ERR_CUDA, "cudaStreamWaitEvent failed");
        if (rk == JB_OK) rk = prepare_molecules(c, cur, c->stream);
        if (rk == JB_OK) rk = query_run_impl(c, t, cur, q, res, cut[k]);       // returns with the slice's kernels done
        if (c->debug >= 2) fprintf(stderr, "[jb]   slice %d searched: %.2f - %.2f ms, %lld hits so far, retries %lld\n", k,
                                   t_run - t_begin, now_ms() - t_begin, res->count, (long long)res->stats.retries);
        if (rk == JB_OK) {
            if (direct && (size_t)res->count <= res->block_cap) {
                rk = result_copy_range(res, copied, (size_t)res->count, c->copy_stream);
                copied = (size_t)res->count;
            } else direct = false;
        }
        if (next.valid()) {                              // always joined: the helper references this frame
            const std::pair<int, std::string> r = next.get();
            if (r.first != JB_OK && rk == JB_OK) rk = fail(r.first, "%s", r.second.c_str());
        }
        rc = rk;
    }
    if (rc == JB_OK) CU(cudaStreamSynchronize(c->copy_stream));
    if (c->debug >= 2) fprintf(stderr, "[jb]   copies done: %.2f ms (direct %d)\n", now_ms() - t_begin, (int)direct);
    if (rc == JB_OK && !direct) {                     // more hits than expected: one plain fetch of everything
        rc = result_alloc_block(res, (size_t)res->count);
        if (rc == JB_OK) rc = result_copy_range(res, 0, (size_t)res->count, c->stream);
        if (rc == JB_OK) CU(cudaStreamSynchronize(c->stream));
    }
    if (rc != JB_OK) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        jb_result_free(res);
        return rc;
    }
    res->stats.hits = (int64_t)res->count;
    c->last_hits = res->count;
    result_finalize(res);
    *out = res;
    return jb_result_fetch(res, hits);
}

static int query_host_once(jb_context *c, const jb_templates *t, const HostSource &src, const jb_query_params *q,
                           jb_result **out, jb_hits *hits);

extern "C" int jb_query_host(jb_context *c, const jb_templates *t, const jb_molecule_batch *b, const jb_query_params *q,
                             jb_result **out, jb_hits *hits)
{
    if (!c || !b || !out || !hits) return fail(JB_ERR_INVALID, "jb_query_host: NULL argument");
    if (b->n_molecules < 0) return fail(JB_ERR_INVALID, "negative molecule count");
    HostSource src;
    src.soa = b;
    src.M = b->n_molecules;
    context_decay(c);
    int rc = query_host_once(c, t, src, q, out, hits);
    if (rc == JB_ERR_NOMEM) {                         // same policy as jb_query_run
        context_trim(c);
        rc = query_host_once(c, t, src, q, out, hits);
    }
    return rc;
}

extern "C" int jb_query_records(jb_context *c, const jb_templates *t, const jb_atom_record *const *molecules,
                                const int64_t *counts, int32_t n_molecules, const jb_query_params *q, jb_result **out,
                                jb_hits *hits)
{
    if (!c || !out || !hits || n_molecules < 0 || (n_molecules > 0 && (!molecules || !counts)))
        return fail(JB_ERR_INVALID, "jb_query_records: bad argument");
    HostSource src;
    src.recs = molecules;
    src.counts = counts;
    src.M = n_molecules;
    src.serial = q && q->hit_order != JB_ORDER_CANONICAL;   // Atom.serial only orders the residue index of the reference
    src.off.assign((size_t)n_molecules + 1, 0);
    for (int m = 0; m < n_molecules; m++) {
        if (counts[m] < 0 || (counts[m] > 0 && !molecules[m])) return fail(JB_ERR_INVALID, "jb_query_records: structure %d", m);
        src.off[m + 1] = src.off[m] + counts[m];
    }
    context_decay(c);
    int rc = query_host_once(c, t, src, q, out, hits);
    if (rc == JB_ERR_NOMEM) {
        context_trim(c);
        rc = query_host_once(c, t, src, q, out, hits);
    }
    return rc;
}

static int query_host_once(jb_context *c, const jb_templates *t, const HostSource &src, const jb_query_params *q,
                           jb_result **out, jb_hits *hits)
{
    *out = nullptr;
    if (!t || !q || t->ctx != c) return fail(JB_ERR_INVALID, "jb_query_host: bad templates/params");
    if (q->ignore_chain < 0 || q->ignore_chain > 2) return fail(JB_ERR_INVALID, "invalid ignore_chain %d", q->ignore_chain);
    if (q->reorder < 0 || q->reorder > 2) return fail(JB_ERR_INVALID, "invalid reorder %d", q->reorder);
    if (q->hit_order < 0 || q->hit_order > 2) return fail(JB_ERR_INVALID, "invalid hit_order %d", q->hit_order);
    CU(cudaSetDevice(c->device));
    if (src.M > 0 && t->T > 0 && t->TA > 0 && c->pipeline_slices != 1) {
        const long long NA = src.offsets()[src.M];
        // auto: slicing costs launches and host syncs per slice (measured on cfg3: 4 slices +1.0 ms,
        // 2 slices -0.25 ms), so only the copies worth hiding are pipelined: 2 slices for a large
        // upload, 4 when the previous call on this context returned a large hit volume
        const long long out_bytes = c->last_hits * (long long)(std::max(t->nmax, 1) * 4 + 136);
        int n_slices = c->pipeline_slices > 1 ? c->pipeline_slices
                       : out_bytes >= (256ll << 20) ? 4 : (NA >= 1000000 && q->hit_order == JB_ORDER_CANONICAL) ? 2 : 1;
        // (the rank tables of the reference hit order are rebuilt per slice: not worth two slices)
        // records mode: the first slice's packing is the only one the engine does not hide -- more, smaller slices
        if (!src.soa && c->pipeline_slices <= 1 && NA >= 1000000)
            n_slices = std::max(n_slices, (int)std::min<long long>(8, std::max<long long>(2, NA / 2000000)));
        n_slices = std::min(n_slices, src.M / 16);
        if (n_slices >= 2) return query_host_pipelined(c, t, src, q, n_slices, out, hits);
    }
    if (!c->scratch_mol) c->scratch_mol = new jb_molecules();
    // grow-only device staging owned by the context
    int rc = stage_source(c, c->scratch_mol, src, 0, src.M, c->stream, c->host_stage[0]);
    if (rc == JB_OK) rc = prepare_molecules(c, c->scratch_mol, c->stream);
    // (no synchronisation here: the run below waits for its own kernels, which follow the copies on the same stream, before
    //  it returns -- the caller's staging is free again by then; a round trip less on the latency path)
    if (rc != JB_OK) return rc;
    rc = query_run_once(c, t, c->scratch_mol, q, out);
    if (rc == JB_OK) c->last_hits = (*out)->count;
    if (rc == JB_OK) rc = jb_result_fetch(*out, hits);
    if (rc != JB_OK) {
        cudaStreamSynchronize(c->stream);             // the uploads may still be reading the caller's staging
        if (*out) { jb_result_free(*out); *out = nullptr; }
    }
    return rc;
}
