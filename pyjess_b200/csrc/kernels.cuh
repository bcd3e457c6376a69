// kernels.cuh -- sm_100a device code of the template-query engine.
//
// One kernel per stage of the reference hot path:
//   k_pack_atoms, k_sort_atoms (or k_sort_keys + a device-wide radix sort + k_sorted_view
//                     for batches with structures > 16,384 atoms): 32-byte atom records and
//                     the x-sorted FP32 view, once per upload (stand in for the per-pair
//                     k-d tree builds, eff:KdTree.c:145-186,343-423)
//   k_classify        atoms x atom-specs -> per-(structure, spec) candidate
//                     lists (replaces TessTemplate_candidates, ResIndex_*,
//                     TessAtom_match: eff:TessTemplate.c:141-201,
//                     eff:ResIndex.c:28-153, eff:TessAtom.c:238-449)
//   k_xindex          32-bucket x index of every candidate list (the range query of
//                     KdTreeQuery_next becomes a table look-up + a short scan)
//   k_shells          squared annulus radii per ordered template-atom pair
//                     (eff:Scanner.c:275-290, eff:Annulus.c:40-67)
//   k_order, k_pairs  per (structure, template): liveness + scan order
//                     (Scanner_reuse, eff:Scanner.c:106-183)
//   k_expand<G>       level-synchronous frontier expansion, G lanes per
//                     partial match (replaces Scanner_next + KdTreeQuery_next
//                     + _Join_ipo/_Annulus_po + TessTemplate_check:
//                     eff:Scanner.c:193-304, eff:KdTree.c:238-296,
//                     eff:Annulus.h:54-68, eff:TessTemplate.c:78-119); run-time level,
//                     used beyond level 12 and for sets with single-atom templates --
//                     the hot levels run k_expand_lv<L,G,ROOT> (expand_lv.cuh)
//   k_superpose       batched FP64 superposition + RMSD filter + compaction
//                     (vendor/jess/src/Super.c:139-538, eff:Jess.c:176-195,
//                     src/pyjess/_jess.pyx:2238-2257)
// hits.cuh orders the hit records, order.cuh rebuilds the reference's iteration order.
//
// Nothing here is a dense contraction: no tensor cores, no TMEM.  All FP64
// decisions use explicit round-to-nearest intrinsics (and the TU is built
// with -fmad=false) because the reference build contains no FMA and the
// accept/reject arithmetic must be reproduced bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jb {

typedef unsigned long long u64;

constexpr int kMaxN = 32;            // JB_MAX_TEMPLATE_ATOMS
constexpr int kStageItems = 64;      // per-warp staging capacity (items)
constexpr int kTupleChunk = 64;      // records of the tuple list a warp reserves with one atomic (expand_grid.cuh)
constexpr int kExpandThreads = 256;
constexpr int kExpandWarps = kExpandThreads / 32;

// ---- device-visible tables -------------------------------------------------

struct SpecTable {            // de-duplicated template-atom specs
    int n_specs;
    const int *code;          // [S]
    const int *name_off;      // [S+1]
    const uint32_t *name_raw; // tokens as given (case sensitive rules)
    const uint32_t *name_up;  // upper-cased (case insensitive rules)
    const int *res_off;       // [S+1]
    const uint32_t *res_up;   // upper-cased residue tokens
};

struct ShellRec;
struct TemplateTable {
    int n_templates;
    const int *atom_off;       // [T+1]
    const int *spec;           // [TA] spec id of each template atom
    const int *resseq;         // [TA]
    const unsigned char *chain;// [TA]
    const double *pos;         // [TA*3]
    const double *weight;      // [TA]
    const long long *dist_off; // [T+1]
    const double *dist;        // n*n per template
    ShellRec *shell;           // same layout as dist: [earlier*n + later]
    double *vol;               // same layout: shell volume proxy hi^3 - lo^3 (scan-order heuristic)
};

struct alignas(32) AtomRec {   // one 32-byte sector per gathered atom (original atom order)
    double x, y, z;
    int resseq;
    int chain;
};

struct alignas(32) MolExt {    // frame of the FP32 records of one structure
    double ox, oy, oz;         // origin (first atom)
    double ext;                // max |coordinate - origin| over the structure
};

// Device image of an uploaded batch.  Two views of the same atoms:
//   atoms[]            original order, FP64, authoritative for every decision
//   srec/sname/sres[]  sorted by x within each structure (k_sort_atoms), FP32 coordinates
//                      relative to the structure origin + original index: conservative
//                      prefilter data only.  Candidate lists inherit the x order.
struct MolTable {
    int n_molecules;           // of the chunk
    const long long *atom_off; // [M+1] (chunk view, absolute atom offsets)
    const AtomRec *atoms;
    const float4 *srec;        // {x-ox, y-oy, z-oz, original index bits}
    const uint32_t *sname, *sres;
    const MolExt *ext;         // [M] chunk view
};

constexpr int kXBuckets = 32;  // x-bucket index of a candidate list (k_xindex)

struct alignas(32) CandHead {  // one candidate list = one (structure, spec): one sector
    long long off;
    int cnt;
    float x0;                  // x of the first (smallest-x) record
    float inv_w;               // kXBuckets / (x range of the list); 0: a single bucket
    int bin0;                  // first BinTab of the list (k_classify reserves ceil(cnt / kBinRecs) of them)
    int pad[2];
};

// bucket of an x coordinate: monotone non-decreasing in x (no FMA: the TU is built with
// -fmad=false, and k_xindex and the scans call this very function)
__device__ __forceinline__ int x_bucket(float x, float x0, float inv_w)
{
    return (int)fminf(fmaxf((x - x0) * inv_w, 0.f), (float)(kXBuckets - 1));
}

// ---- the cell grid of a candidate list (k_binindex, expand_grid.cuh) ---------------------------
// A list is cut into BINS of kBinRecs consecutive records (x-sorted, so a bin is an x slab).  Every
// bin carries a uniform grid of kMaskBuckets cells per axis over its own bounding box, stored
// separably: per axis and cell b the BIT MASK of the bin's records whose cell index is <= b.  The
// records inside an axis-aligned box are then
//     AND over the three axes of  upto[a][cell(hi_a)] & ~upto[a][cell(lo_a) - 1]
// -- six 16-byte loads and a few logic operations per (partial match, bin) instead of one load and six
// compares per RECORD (the k-d tree walk of eff:KdTree.c:238-296 becomes three slab look-ups).  The
// same monotone cell function bins the records and the query bounds, so the result is a superset
// of the records inside the box; the exact FP64 predicate still decides.
constexpr int kBinRecs = 128;          // (64-record bins with 8-byte masks measured the same on cfg 3 and cfg 4, gpurun_out/r2c)
constexpr int kMaskWords = kBinRecs / 32;
constexpr int kMaskBuckets = 32;

struct alignas(kBinRecs / 8) BinMask { unsigned w[kMaskWords]; };

struct alignas(32) BinTab {            // 1568 bytes
    float2 frame[4];                   // per axis: (smallest coordinate of the bin, cells per unit length); [3] unused
    BinMask upto[3][kMaskBuckets];
};

__device__ __forceinline__ BinMask ldg_mask(const BinMask *p)
{
    BinMask r;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
    return r;
}

__device__ __forceinline__ int axis_cell(float v, float2 frame)
{
    return (int)fminf(fmaxf((v - frame.x) * frame.y, 0.f), (float)(kMaskBuckets - 1));
}

struct CandTable {
    CandHead *head;            // [M*S]
    unsigned short *xtab;      // [M*S][kXBuckets]: first record whose x bucket is >= b (lists <= 65535 records)
    float4 *rec;               // x-sorted copies of srec entries: FP32 is ONLY a conservative prefilter
    long long cap;             // capacity (records)
    BinTab *bins;              // cell grids, one per bin (k_binindex)
    int *binlist;              // bin -> list
    long long bin_cap;         // capacity (bins)
};

struct alignas(32) ShellRec {  // constraint of an ORDERED template-atom pair (earlier i, later k)
    double lo2, hi2;           // squared shell radii (eff:Scanner.c:275-290, eff:Annulus.c:40-67)
    int flags;                 // bit0: same template chain, bit1: same template residue number
    float rb;                  // outer radius rounded UP to FP32 (join-box prefilter)
    int pad[2];
};

struct alignas(16) PairRec {   // one live (structure, template) pair, 128 bytes
    int m;                     // structure index within the chunk
    int t;                     // template index
    int ta0;                   // first template atom
    int n;                     // template size
    long long abase;           // first atom of the structure
    long long shell0;          // first ShellRec of the template
    unsigned short spec[kMaxN];// spec id of the atom scanned at level l
    unsigned char order[kMaxN];// scan order (eff:Scanner.c:88-99,160-174 or the engine's own)
};

struct Counters {              // device counters of one chunk attempt
    u64 cand_total;            // records reserved in CandTable
    u64 bin_total;             // bins reserved in CandTable
    u64 live_pairs;
    u64 cand_atoms;            // sum_k |C_k| over live pairs
    u64 cand_tests;            // sum over partial matches of |C_level| (algorithm independent)
    u64 scan_tests;            // FP32 box tests actually executed (after x-range narrowing)
    u64 box_pass;              // ... that survived and took the exact FP64 predicate
    u64 tuples;                // full tuples produced
    u64 tuple_slots;           // records reserved in the tuple list: the level kernels reserve kTupleChunk at a time per warp
                               // and mark the records a chunk leaves unused as holes (pair = -1); >= tuples
    u64 tuple_ints;            // sum of their sizes
    u64 hits;                  // hits written (absolute, across chunks)
    u64 nan_dropped;
    u64 level[kMaxN + 1];      // level[l] = items with l atoms (l>=1)
    u64 sched[kMaxN + 1];      // sched[l] = next unclaimed item (pair in root mode) of the level-l sweep: warps claim work here
    u64 tail_level;            // != 0: the frontier of that level was small enough to be finished depth first by
                               // k_expand_tail (expand_dfs.cuh); the level kernels of that and of deeper levels return at once
};

// ---- tokens ------------------------------------------------------------------

__device__ __forceinline__ uint32_t upper4(uint32_t t)
{
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t c = (t >> (8 * i)) & 0xffu;
        if (c >= 'a' && c <= 'z') c &= ~0x20u;
        r |= c << (8 * i);
    }
    return r;
}
__device__ __forceinline__ uint32_t byte_of(uint32_t t, int i) { return (t >> (8 * i)) & 0xffu; }

__host__ __device__ constexpr uint32_t tok(char a, char b, char c, char d)
{
    return (uint32_t)(unsigned char)a | ((uint32_t)(unsigned char)b << 8) | ((uint32_t)(unsigned char)c << 16) |
           ((uint32_t)(unsigned char)d << 24);
}

// TessAtom_match (eff:TessAtom.c:324-449) for one atom against one spec.
// name = raw token, name_up/res_up = upper-cased tokens of the atom.
__device__ __forceinline__ bool spec_match(const SpecTable &S, int s, uint32_t name, uint32_t name_up, uint32_t res_up)
{
    const int code = __ldg(S.code + s);
    const int n0 = __ldg(S.name_off + s), n1 = __ldg(S.name_off + s + 1);
    const int r0 = __ldg(S.res_off + s), r1 = __ldg(S.res_off + s + 1);
    const bool under = byte_of(name, 0) == '_';
    const bool carbon = under && byte_of(name, 1) == 'C';      // eff:TessAtom.c:238-241
    const bool hydrogen = under && byte_of(name, 1) == 'H';    // :244-247
    const bool mainchain = name_up == tok('_', 'C', 'A', '_') || name_up == tok('_', 'N', '_', '_') ||
                           name_up == tok('_', 'O', '_', '_'); // :269-276
    bool ok;
    switch (code) {
    case -1: case 0: case 100: {                               // matchName :278-288
        ok = false;
        for (int k = n0; k < n1; k++) ok |= (name_up == __ldg(S.name_up + k));
        break;
    }
    case 1: ok = !carbon && !hydrogen && !mainchain; break;
    case 2: ok = !carbon && !hydrogen; break;
    case 3: case 103: {                                        // matchType :302-322 (case sensitive)
        ok = false;
        for (int k = n0; k < n1; k++) {
            uint32_t t = __ldg(S.name_raw + k);
            ok |= under ? (byte_of(name, 1) == byte_of(t, 1))
                        : (byte_of(name, 0) == byte_of(t, 0) && byte_of(name, 1) == byte_of(t, 1));
        }
        if (code == 3) ok = ok && !hydrogen;
        break;
    }
    case 4: ok = !carbon && !hydrogen && mainchain; break;
    case 5: ok = mainchain && !hydrogen; break;
    case 6: ok = !mainchain && !hydrogen; break;
    case 7: ok = !hydrogen; break;
    case 8: {                                                  // isInSamePosition :249-266
        ok = false;
        for (int k = n0; k < n1; k++) {
            uint32_t t = __ldg(S.name_raw + k);
            ok |= under ? (byte_of(name, 2) == byte_of(t, 2))
                        : (byte_of(name, 1) == byte_of(t, 1) && byte_of(name, 2) == byte_of(t, 2));
        }
        ok = ok && !hydrogen;
        break;
    }
    case 101: ok = !carbon && !mainchain; break;
    case 102: ok = !carbon; break;
    case 104: ok = !carbon && mainchain; break;                // falls through into 105
    case 105: ok = mainchain; break;
    case 106: ok = !mainchain; break;
    case 107: ok = true; break;
    default: ok = false; break;
    }
    if (ok && code < 100) {                                    // matchResName :290-300
        bool r = false;
        for (int k = r0; k < r1; k++) r |= (res_up == __ldg(S.res_up + k));
        ok = r;
    }
    return ok;
}

// ---- K0: atom records ---------------------------------------------------------
// Runs once per upload: SoA host image -> one 32-byte record per atom, so that every
// later gather of an atom costs exactly one sector.
__global__ void __launch_bounds__(256) k_pack_atoms(long long n, const double *x, const double *y, const double *z,
                                                    const int *resseq, const unsigned char *chain, AtomRec *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    AtomRec r;
    r.x = x[i]; r.y = y[i]; r.z = z[i];
    r.resseq = resseq[i];
    r.chain = (int)chain[i];
    out[i] = r;
}

// ---- K0b: per-structure frame + x-sorted FP32 view (once per upload) ----------------
// One CTA per structure: origin = first atom, extent = max |coordinate - origin|
// (bounds the FP32 rounding of the relative coordinates), then a bitonic sort of
// (orderable bits of float(x - ox), original index) keys -- in shared memory when
// the padded structure fits, in a global scratch area otherwise.  Replaces the k-d
// tree build of the reference (eff:KdTree.c:145-186,343-423): every candidate list
// later cut out of this view is x-sorted, so a range query is a lower-bound search
// plus a short scan.
__device__ __forceinline__ uint32_t float_orderable(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kSortThreads = 512;

struct RawAtoms {              // SoA upload image; x == nullptr: the 32-byte records exist already (k_pack_atoms)
    const double *x, *y, *z;
    const int *resseq;
    const unsigned char *chain;
};

__global__ void __launch_bounds__(kSortThreads) k_sort_atoms(const long long *atom_off, AtomRec *atoms,
                                                             const uint32_t *name, const uint32_t *res, MolExt *ext,
                                                             float4 *srec, uint32_t *sname, uint32_t *sres,
                                                             u64 *scratch, int smem_keys, RawAtoms raw)
{
    extern __shared__ __align__(16) unsigned char sort_smem[];
    __shared__ double sh_ext[kSortThreads / 32];
    const int m = blockIdx.x;
    const long long a0 = atom_off[m], a1 = atom_off[m + 1];
    const int na = (int)(a1 - a0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (raw.x) {               // small batches: K0 fused in (one launch less on the latency path)
        for (int i = tid; i < na; i += kSortThreads) {
            AtomRec r;
            r.x = raw.x[a0 + i]; r.y = raw.y[a0 + i]; r.z = raw.z[a0 + i];
            r.resseq = raw.resseq[a0 + i];
            r.chain = (int)raw.chain[a0 + i];
            atoms[a0 + i] = r;
        }
        __syncthreads();       // the records of this structure are read back below, by other threads of the CTA
    }
    double ox = 0.0, oy = 0.0, oz = 0.0, e = 0.0;
    if (na > 0) { const AtomRec o = atoms[a0]; ox = o.x; oy = o.y; oz = o.z; }
    for (int i = tid; i < na; i += kSortThreads) {
        const AtomRec ar = atoms[a0 + i];
        e = fmax(e, fmax(fabs(__dsub_rn(ar.x, ox)), fmax(fabs(__dsub_rn(ar.y, oy)), fabs(__dsub_rn(ar.z, oz)))));
    }
    for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
    if (lane == 0) sh_ext[w] = e;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int i = 0; i < kSortThreads / 32; i++) v = fmax(v, sh_ext[i]);
        MolExt me; me.ox = ox; me.oy = oy; me.oz = oz; me.ext = v;
        ext[m] = me;
    }
    if (na == 0) return;
    int npad = 1;
    while (npad < na) npad <<= 1;
    u64 *keys = npad <= smem_keys ? reinterpret_cast<u64 *>(sort_smem) : scratch + 2 * a0;
    for (int i = tid; i < npad; i += kSortThreads) {
        u64 k = ~0ull;
        if (i < na) {
            const float fx = __double2float_rn(__dsub_rn(atoms[a0 + i].x, ox));
            k = ((u64)float_orderable(fx) << 32) | (u64)(uint32_t)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += kSortThreads) {
                const int p = i ^ j;
                if (p > i) {
                    const u64 x = keys[i], y = keys[p];
                    if ((x > y) == ((i & k) == 0)) { keys[i] = y; keys[p] = x; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < na; i += kSortThreads) {
        const int a = (int)(uint32_t)keys[i];
        const AtomRec ar = atoms[a0 + a];
        srec[a0 + i] = make_float4(__double2float_rn(__dsub_rn(ar.x, ox)), __double2float_rn(__dsub_rn(ar.y, oy)),
                                   __double2float_rn(__dsub_rn(ar.z, oz)), __int_as_float(a));
        sname[a0 + i] = name[a0 + a];
        sres[a0 + i] = res[a0 + a];
    }
}

// ---- K0c: the same view for batches holding very large structures --------------------------
// k_sort_atoms sorts one structure per CTA; beyond 16,384 atoms its keys leave shared memory and
// a single CTA runs ~130 bitonic passes over global memory (5 ms for the 58,870-atom structure of
// BASELINE config 2).  Such batches are sorted by ONE device-wide radix sort over the keys
// (structure | orderable x): k_sort_keys (frame + keys), cub::DeviceRadixSort, k_sorted_view.
__global__ void __launch_bounds__(kSortThreads) k_sort_keys(const long long *atom_off, const AtomRec *atoms, MolExt *ext,
                                                            u64 *key, unsigned *idx)
{
    __shared__ double sh_ext[kSortThreads / 32];
    const int m = blockIdx.x;
    const long long a0 = atom_off[m], a1 = atom_off[m + 1];
    const int na = (int)(a1 - a0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double ox = 0.0, oy = 0.0, oz = 0.0, e = 0.0;
    if (na > 0) { const AtomRec o = atoms[a0]; ox = o.x; oy = o.y; oz = o.z; }
    for (int i = tid; i < na; i += kSortThreads) {
        const AtomRec ar = atoms[a0 + i];
        e = fmax(e, fmax(fabs(__dsub_rn(ar.x, ox)), fmax(fabs(__dsub_rn(ar.y, oy)), fabs(__dsub_rn(ar.z, oz)))));
        const float fx = __double2float_rn(__dsub_rn(ar.x, ox));
        key[a0 + i] = ((u64)(unsigned)m << 32) | (u64)float_orderable(fx);
        idx[a0 + i] = (unsigned)(a0 + i);
    }
    for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
    if (lane == 0) sh_ext[w] = e;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int i = 0; i < kSortThreads / 32; i++) v = fmax(v, sh_ext[i]);
        MolExt me; me.ox = ox; me.oy = oy; me.oz = oz; me.ext = v;
        ext[m] = me;
    }
}

__global__ void __launch_bounds__(256) k_sorted_view(long long n, const u64 *skey, const unsigned *sidx,
                                                     const long long *atom_off, const AtomRec *atoms, const uint32_t *name,
                                                     const uint32_t *res, const MolExt *ext, float4 *srec, uint32_t *sname,
                                                     uint32_t *sres)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int m = (int)(skey[j] >> 32);
    const long long g = (long long)sidx[j];
    const AtomRec ar = atoms[g];
    const MolExt me = ext[m];
    srec[j] = make_float4(__double2float_rn(__dsub_rn(ar.x, me.ox)), __double2float_rn(__dsub_rn(ar.y, me.oy)),
                          __double2float_rn(__dsub_rn(ar.z, me.oz)), __int_as_float((int)(g - atom_off[m])));
    sname[j] = name[g];
    sres[j] = res[g];
}

// ---- K1: candidate lists ------------------------------------------------------
// One CTA per structure, atoms visited in x-sorted order.  Warp w owns a contiguous
// slice; pass 1 counts matches per (warp, spec) in shared memory, a scan over warps
// gives every warp its write offset, pass 2 re-evaluates and copies the FP32 records,
// so every list is x-sorted and deterministic.  List placement in the record array is
// reserved with one atomicAdd per (structure, spec).
//
// TessAtom_match only looks at (atom name, residue name), and a structure has a few
// hundred distinct such pairs at most: the CTA first collects them in a shared-memory
// hash set (kClassSlots slots), evaluates every (class, spec) pair ONCE into a bit mask,
// and the two passes then visit, per 32 atoms, only the specs that some atom of the group
// matches (union of the lanes' masks) instead of all of them: 30x fewer spec_match
// evaluations on BASELINE config 3.  A structure with more classes than the set holds, or a
// template set whose masks do not fit (class_words == 0), takes the plain loop over all specs.
constexpr int kClassSlots = 1024;
constexpr int kClassMaxFill = 768;

__device__ __forceinline__ unsigned class_hash(uint32_t name, uint32_t res_up)
{
    return ((name * 0x9E3779B1u) ^ (res_up * 0x85EBCA6Bu) ^ (name >> 15)) & (kClassSlots - 1);
}

__device__ __forceinline__ void xindex_list(const long long L, const int lane, const CandTable &C);

// 32 x 32 bit transpose across the lanes of a warp: lane b returns the word whose bit a is bit b of lane a's `x`
// (five butterfly exchanges; checked against the definition in tests/test_bit_transpose.py's model of this code).
__device__ __forceinline__ unsigned warp_transpose32(unsigned x, const int lane)
{
    unsigned m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, j);
        if ((lane & j) == 0) x ^= (((x >> j) ^ y) & m) << j;       // keep my low halves, take the partner's into the high ones
        else x ^= ((y >> j) ^ x) & m;                               // keep my high halves, take the partner's into the low ones
        m ^= m << (j >> 1);
    }
    return x;
}

__global__ void __launch_bounds__(1024) k_classify(MolTable M, SpecTable S, CandTable C, Counters *ctr, int class_words)
{
    extern __shared__ int sh_cnt[];                 // [warps][S], then the class set and its masks
    __shared__ int sh_fill;
    const int m = blockIdx.x;
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ns = S.n_specs;
    const long long a0 = M.atom_off[m], a1 = M.atom_off[m + 1];
    const int na = (int)(a1 - a0);
    u64 *ckey = reinterpret_cast<u64 *>(sh_cnt + ((nw * ns + 1) & ~1));          // [kClassSlots]
    unsigned *cmask = reinterpret_cast<unsigned *>(ckey + kClassSlots);           // [kClassSlots][class_words]
    long long *sh_off = reinterpret_cast<long long *>(cmask + (size_t)kClassSlots * class_words);   // [ns] first record of each list (fast path)
    unsigned *tb = reinterpret_cast<unsigned *>(sh_off + ns) + w * 32;            // [warps][32] transposed match masks of the current group
    const unsigned lt = (1u << lane) - 1;
    constexpr u64 EMPTY = ~0ull;
    bool fast = class_words > 0;
    if (fast) {
        for (int i = threadIdx.x; i < kClassSlots; i += blockDim.x) ckey[i] = EMPTY;
        if (threadIdx.x == 0) sh_fill = 0;
        __syncthreads();
        for (int a = threadIdx.x; a < na; a += blockDim.x) {
            const uint32_t name = M.sname[a0 + a], res_up = upper4(M.sres[a0 + a]);
            const u64 key = ((u64)name << 32) | res_up;
            if (key == EMPTY) { atomicAdd(&sh_fill, kClassSlots); continue; }     // (0xFF bytes only) -> plain loop
            unsigned h = class_hash(name, res_up);
            for (int probe = 0; probe < kClassSlots; probe++) {
                const u64 cur = ckey[h];
                if (cur == key) break;
                if (cur == EMPTY) {
                    const u64 old = atomicCAS(reinterpret_cast<unsigned long long *>(&ckey[h]), EMPTY, key);
                    if (old == EMPTY) { atomicAdd(&sh_fill, 1); break; }
                    if (old == key) break;
                }
                h = (h + 1) & (kClassSlots - 1);
                if (*(volatile int *)&sh_fill > kClassMaxFill) break;
            }
        }
        __syncthreads();
        fast = sh_fill <= kClassMaxFill;                                          // uniform over the CTA
        if (fast) {
            for (int i = threadIdx.x; i < kClassSlots * class_words; i += blockDim.x) cmask[i] = 0u;
            __syncthreads();
            // one (class, spec) evaluation per thread-iteration; a warp takes a class and walks its specs 32 at a time
            // (slot by slot: the flat loop over (slot, word) jobs spent 29 % of the kernel on its division and on empty slots)
            for (int slot = w; slot < kClassSlots; slot += nw) {
                const u64 key = ckey[slot];
                if (key == EMPTY) continue;
                const uint32_t name = (uint32_t)(key >> 32), name_up = upper4(name), res_up = (uint32_t)key;
                for (int word = 0; word < class_words; word++) {
                    const int sp = word * 32 + lane;
                    const bool ok = sp < ns && spec_match(S, sp, name, name_up, res_up);
                    const unsigned bits = __ballot_sync(0xffffffffu, ok);
                    if (lane == 0) cmask[slot * class_words + word] = bits;
                }
            }
        }
        __syncthreads();
    }
    // slice boundaries in multiples of 32 atoms
    const int chunks = (na + 31) >> 5;
    const int c_lo = (int)(((long long)chunks * w) / nw), c_hi = (int)(((long long)chunks * (w + 1)) / nw);
    int *mine = sh_cnt + w * ns;
    for (int s = lane; s < ns; s += 32) mine[s] = 0;
    __syncwarp();
    for (int pass = 0; pass < 2; pass++) {
        for (int c = c_lo; c < c_hi; c++) {
            const int a = (c << 5) + lane;
            const bool valid = a < na;
            uint32_t name = 0, name_up = 0, res_up = 0;
            float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                name = M.sname[a0 + a];
                name_up = upper4(name);
                res_up = upper4(M.sres[a0 + a]);
                if (pass == 1) rec = M.srec[a0 + a];
            }
            // one step of either loop: lanes with `ok` match spec s
            auto visit = [&](int s, bool ok) {
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                if (mask == 0) return;
                if (pass == 0) {
                    if (lane == 0) mine[s] += __popc(mask);
                } else {
                    const int base = mine[s];
                    if (ok) {
                        const long long dst = C.head[(long long)m * ns + s].off + base + __popc(mask & ((1u << lane) - 1));
                        if (dst < C.cap) C.rec[dst] = rec;
                    }
                    __syncwarp();
                    if (lane == 0) mine[s] = base + __popc(mask);
                }
                __syncwarp();
            };
            if (fast) {
                const unsigned *mymask = nullptr;
                if (valid) {
                    const u64 key = ((u64)name << 32) | res_up;
                    unsigned h = class_hash(name, res_up);
                    while (ckey[h] != key) h = (h + 1) & (kClassSlots - 1);
                    mymask = cmask + h * class_words;
                }
                // Per 32 atoms and 32 specs ONE bit transpose instead of one ballot per spec some atom matches (40-100 of
                // them, most matched by one or two lanes: 88 warp instructions per atom, 0.46 ms per cfg 3 step, 15 % of a
                // cfg 4 step): lane b then owns spec 32 word + b and holds the lanes that match it -- all 32 counters move
                // in one instruction, and an atom finds its slot in each of ITS specs' lists (a handful) from that table.
                for (int word = 0; word < class_words; word++) {
                    const unsigned bits = valid ? mymask[word] : 0u;
                    if (__reduce_or_sync(0xffffffffu, bits) == 0u) continue;          // warp uniform
                    const unsigned bal = warp_transpose32(bits, lane);
                    const int sp = word * 32 + lane;                                  // (bal == 0 beyond the last spec)
                    if (pass == 1) {
                        tb[lane] = bal;
                        __syncwarp();
                        unsigned mybits = bits;
                        while (mybits) {
                            const int bit = __ffs(mybits) - 1;
                            mybits &= mybits - 1;
                            const int s = word * 32 + bit;
                            const long long dst = sh_off[s] + mine[s] + __popc(tb[bit] & lt);
                            if (dst < C.cap) C.rec[dst] = rec;
                        }
                        __syncwarp();
                    }
                    if (bal) mine[sp] += __popc(bal);
                    __syncwarp();
                }
            } else {
                for (int s = 0; s < ns; s++) visit(s, valid && spec_match(S, s, name, name_up, res_up));
            }
        }
        if (pass == 0) {
            __syncthreads();
            // exclusive scan over warps, reserve the list, publish count/offset
            for (int s = threadIdx.x; s < ns; s += blockDim.x) {
                int run = 0;
                for (int v = 0; v < nw; v++) {
                    const int c = sh_cnt[v * ns + s];
                    sh_cnt[v * ns + s] = run;
                    run += c;
                }
                CandHead h; h.off = 0; h.cnt = run; h.x0 = 0.f; h.inv_w = 0.f; h.bin0 = 0; h.pad[0] = h.pad[1] = 0;
                if (run > 0) {
                    h.off = (long long)atomicAdd(&ctr->cand_total, (u64)run);
                    const int nb = (run + kBinRecs - 1) / kBinRecs;
                    const long long b0 = (long long)atomicAdd(&ctr->bin_total, (u64)nb);
                    h.bin0 = (int)min(b0, 0x7fffffffll);
                    for (int b = 0; b < nb; b++)
                        if (b0 + b < C.bin_cap) C.binlist[b0 + b] = m * ns + s;
                }
                C.head[(long long)m * ns + s] = h;
                if (class_words > 0) sh_off[s] = h.off;
            }
            __syncthreads();
        }
    }
    // K1b fused in: the x-bucket index of this structure's lists (its records were written by this CTA)
    __syncthreads();
    for (int s = w; s < ns; s += nw) xindex_list((long long)m * ns + s, lane, C);
}

// ---- K1b: x-bucket index of every candidate list ---------------------------------------
// The lists are x-sorted; the frontier sweep needs, per partial match, the record range whose
// x lies inside the join box.  Instead of two binary searches (5-6 DEPENDENT scattered loads
// each, 9 % of the sweep's stall samples in profiles/r1g) every list gets a table of
// kXBuckets uniform x buckets: tab[b] = first record with bucket >= b.  One warp per list.
__device__ __forceinline__ void xindex_list(const long long L, const int lane, const CandTable &C)
{
    CandHead hd = C.head[L];
    unsigned short *tab = C.xtab + L * kXBuckets;
    if (hd.cnt <= 0 || hd.off + hd.cnt > C.cap) return;        // empty, or the chunk is being re-run larger
    const float4 *rec = C.rec + hd.off;
    const float x0 = rec[0].x, x1 = rec[hd.cnt - 1].x;
    float inv_w = 0.f;
    if (hd.cnt <= 65535 && x1 > x0) inv_w = (float)kXBuckets / (x1 - x0);
    if (!(inv_w > 0.f) || !(inv_w < 3.0e38f)) inv_w = 0.f;      // NaN / overflow: one bucket
    if (lane == 0) { hd.x0 = x0; hd.inv_w = inv_w; C.head[L] = hd; }
    if (hd.cnt > 65535) return;                                // scans fall back to the whole list
    for (int p = lane; p < hd.cnt; p += 32) {
        const int b = x_bucket(rec[p].x, x0, inv_w);
        const int bp = p > 0 ? x_bucket(rec[p - 1].x, x0, inv_w) : -1;
        for (int q = bp + 1; q <= b; q++) tab[q] = (unsigned short)p;
    }
    const int bl = x_bucket(x1, x0, inv_w);
    for (int q = bl + 1 + lane; q < kXBuckets; q += 32) tab[q] = (unsigned short)hd.cnt;
}

__global__ void __launch_bounds__(256) k_xindex(long long n_lists, CandTable C)
{
    const long long L = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (L >= n_lists) return;
    xindex_list(L, threadIdx.x & 31, C);
}

// ---- K1c: cell grid of every bin --------------------------------------------------------
// One warp per bin: bounding box (warp min / max), cell of every record along the three axes,
// 32 ballots per axis and 32 records of the bin.  Records with a NaN coordinate enter no mask (the box
// test they replace -- r.x >= lo && r.x <= hi ... -- never accepted them either).
// (Round 2 tried ONE word per cell for bins of at most 32 records, the 32 cells of an axis in one 128-byte line -- 3 lines
//  per (partial match, bin) instead of 6: 7.67 -> 7.74 ms on cfg 3, 12.83 -> 12.82 ms on cfg4shard/4, gpurun_out/r3j; dropped.)
__global__ void __launch_bounds__(256) k_binindex(CandTable C, const Counters *ctr)
{
    const int lane = threadIdx.x & 31;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long n_bins = (long long)ctr->bin_total;
    if (n_bins > C.bin_cap) n_bins = C.bin_cap;                 // the chunk is being re-run larger
    const float INF = __int_as_float(0x7f800000);
    for (long long bin = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; bin < n_bins; bin += n_warps) {
        const CandHead hd = C.head[C.binlist[bin]];
        BinTab *T = C.bins + bin;
        const int first = (int)(bin - hd.bin0) * kBinRecs;
        if (hd.cnt <= 0 || hd.off + hd.cnt > C.cap || first < 0 || first >= hd.cnt) {
            // a list cut by the end of the record array (the chunk is being re-run larger): an EMPTY grid, so
            // that the sweep of this attempt never follows a stale mask to a record that was not written
            BinMask none;
#pragma unroll
            for (int h = 0; h < kMaskWords; h++) none.w[h] = 0u;
#pragma unroll
            for (int a = 0; a < 3; a++) T->upto[a][lane] = none;
            if (lane < 4) T->frame[lane] = make_float2(0.f, 0.f);
            continue;
        }
        const int n = min(hd.cnt - first, kBinRecs);
        const float4 *rec = C.rec + hd.off + first;
        float4 r[kMaskWords];
        bool ok[kMaskWords];
        float lo[3] = { INF, INF, INF }, hi[3] = { -INF, -INF, -INF };
#pragma unroll
        for (int h = 0; h < kMaskWords; h++) {
            const int p = lane + 32 * h;
            r[h] = p < n ? rec[p] : make_float4(0.f, 0.f, 0.f, 0.f);
            ok[h] = p < n && r[h].x == r[h].x && r[h].y == r[h].y && r[h].z == r[h].z;
            if (ok[h]) {
                lo[0] = fminf(lo[0], r[h].x); hi[0] = fmaxf(hi[0], r[h].x);
                lo[1] = fminf(lo[1], r[h].y); hi[1] = fmaxf(hi[1], r[h].y);
                lo[2] = fminf(lo[2], r[h].z); hi[2] = fmaxf(hi[2], r[h].z);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            }
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float inv = 0.f;
            if (hi[a] > lo[a]) inv = (float)kMaskBuckets / (hi[a] - lo[a]);
            if (!(inv > 0.f) || !(inv < 3.0e38f)) inv = 0.f;    // NaN / overflow / a single point: one cell
            const float2 fr = make_float2(lo[a] <= hi[a] ? lo[a] : 0.f, inv);
            int cell[kMaskWords];
#pragma unroll
            for (int h = 0; h < kMaskWords; h++) {
                const float v = a == 0 ? r[h].x : (a == 1 ? r[h].y : r[h].z);
                cell[h] = ok[h] ? axis_cell(v, fr) : kMaskBuckets;
            }
            BinMask mine;
#pragma unroll
            for (int h = 0; h < kMaskWords; h++) mine.w[h] = 0u;
            for (int b = 0; b < kMaskBuckets; b++) {
#pragma unroll
                for (int h = 0; h < kMaskWords; h++) {
                    const unsigned mb = __ballot_sync(0xffffffffu, cell[h] <= b);
                    if (lane == b) mine.w[h] = mb;
                }
            }
            T->upto[a][lane] = mine;
            if (lane == 0) T->frame[a] = fr;
        }
        if (lane == 0) T->frame[3] = make_float2(0.f, 0.f);
    }
}

// ---- K: shell radii ------------------------------------------------------------
// shell[earlier i][later k]: the ORDERED pair matters because the dynamic
// threshold is summed left to right as (cutoff + w[i]) + w[k] (eff:Scanner.c:279).
__global__ void k_shells(TemplateTable T, double cutoff, double max_dyn)
{
    const int t = blockIdx.x;
    const int ta0 = T.atom_off[t], n = T.atom_off[t + 1] - ta0;
    const long long d0 = T.dist_off[t];
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, k = e % n;
        const double d = T.dist[d0 + e];
        double dyn = __dadd_rn(__dadd_rn(cutoff, T.weight[ta0 + i]), T.weight[ta0 + k]);
        if (dyn > max_dyn) dyn = max_dyn;
        double a = __dsub_rn(d, dyn), b = __dadd_rn(d, dyn);
        if (a < 0.5) a = 0.5;
        if (b < a) { const double tmp = b; b = a; a = tmp; }   // eff:Annulus.c:46-51
        if (a < 0.0) a = 0.0;
        if (b < 0.0) b = 0.0;
        ShellRec r;
        r.lo2 = __dmul_rn(a, a);
        r.hi2 = __dmul_rn(b, b);
        r.flags = ((T.chain[ta0 + i] == T.chain[ta0 + k]) ? 1 : 0) | ((T.resseq[ta0 + i] == T.resseq[ta0 + k]) ? 2 : 0);
        // outer radius for the FP32 join box: rounded up, with a relative margin far above 1 ulp
        r.rb = __double2float_ru(__dmul_ru(__dsqrt_ru(r.hi2), 1.000000001));
        r.pad[0] = r.pad[1] = 0;
        T.shell[d0 + e] = r;
        // scan-order heuristic (reorder == 2): volume proxy hi^3 - lo^3, exactly as the oracle computes it
        T.vol[d0 + e] = __dsub_rn(__dmul_rn(r.hi2, __dsqrt_rn(r.hi2)), __dmul_rn(r.lo2, __dsqrt_rn(r.lo2)));
    }
}

// ---- K: scan order (reorder == 2) ---------------------------------------------------
// The engine's own scan order: an expected fan-out model.  A candidate of template atom
// k lands in the shell around a placed atom with probability ~ shell volume / structure
// volume, so placing k multiplies the frontier by f(k) = |C_k| * 4.19 * min_j vol(j,k) / V
// (V = 19 A^3 per atom); atoms sharing a template residue with a placed atom can only
// match inside that residue (TessTemplate_check), never prune, and go last, smallest set
// first.  The first atom minimises |C_c| * (1 + best follow-up fan-out).  On cfg 3 this
// cuts the accepted partial matches by 29 % against the reference's argsort(|C_k|).
// One CTA = one template x 128 structures, so the n x n volume table is staged once in
// shared memory and read as warp-uniform broadcasts.  The arithmetic mirrors
// oracle/jess_oracle.c operation by operation (explicit RN ops): same order, same counters.
// (Round 2 tried the per-thread arrays -- counts, running minimum volumes, the order -- in dynamic shared memory instead of
//  local memory, which removes 436 MB of DRAM write-back per cfg 3 step: 0.50 -> 0.67 ms, slower -- the kernel is bound by
//  the issue slots of its O(n^2) loops, not by the local-memory traffic; gpurun_out/r3g.)
__global__ void __launch_bounds__(128) k_order(int n_mol, TemplateTable T, int n_specs, const CandHead *heads,
                                               const long long *atom_off, unsigned char *ordbuf)
{
    __shared__ double s_vol[kMaxN * kMaxN];
    __shared__ unsigned char s_same[kMaxN * kMaxN];
    __shared__ int s_spec[kMaxN];
    const int t = blockIdx.x;
    const int ta0 = T.atom_off[t], n = T.atom_off[t + 1] - ta0;
    const long long d0 = T.dist_off[t];
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        s_vol[e] = T.vol[d0 + e];
        s_same[e] = (T.shell[d0 + e].flags & 3) == 3;
    }
    for (int k = threadIdx.x; k < n; k += blockDim.x) s_spec[k] = T.spec[ta0 + k];
    __syncthreads();
    const int m = blockIdx.y * blockDim.x + threadIdx.x;
    if (m >= n_mol || n < 1) return;
    int cnts[kMaxN];
    bool live = true;
    for (int k = 0; k < n; k++) {
        cnts[k] = heads[(long long)m * n_specs + s_spec[k]].cnt;
        live = live && cnts[k] > 0;
    }
    if (!live) return;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const double V = __dmul_rn(19.0, (double)(atom_off[m + 1] - atom_off[m]));
    // all fan-outs share the divisor V: candidates are compared by their NUMERATORS
    // (|C_k| * 4.19) * vol and only the winner is divided (division by a positive constant is
    // monotone, so min over quotients == quotient of the min numerator, bit for bit)
    double c419[kMaxN];
    for (int k = 0; k < n; k++) c419[k] = __dmul_rn((double)cnts[k], 4.19);
    int first = 0;
    double first_score = inf;
    for (int c = 0; c < n; c++) {
        double mnum = inf;
        for (int k = 0; k < n; k++) {
            if (k == c || s_same[c * n + k]) continue;
            const double num = __dmul_rn(c419[k], s_vol[c * n + k]);
            if (num < mnum) mnum = num;
        }
        const double mf = mnum == inf ? 1.0 : __ddiv_rn(mnum, V);
        const double sc = __dmul_rn((double)cnts[c], __dadd_rn(1.0, mf));
        if (sc < first_score) { first_score = sc; first = c; }
    }
    alignas(16) unsigned char ord[kMaxN];
    double minv[kMaxN];
    unsigned used = 1u << first, same = 0;
    for (int k = 0; k < kMaxN; k++) { minv[k] = inf; ord[k] = 0; }
    ord[0] = (unsigned char)first;
    int last = first;
    for (int l = 1; l < n; l++) {
        // best atom of each class: outside the placed residues by numerator, inside by 1e3 + |C_k|
        int bn = -1, bs = -1;
        double bnum = 0.0, bsf = 0.0;
        for (int k = 0; k < n; k++) {
            if (used & (1u << k)) continue;
            const double v = s_vol[last * n + k];
            if (v < minv[k]) minv[k] = v;
            if (s_same[last * n + k]) same |= 1u << k;
            if (same & (1u << k)) {
                const double f = __dadd_rn(1e3, (double)cnts[k]);
                if (bs < 0 || f < bsf) { bs = k; bsf = f; }
            } else {
                const double num = __dmul_rn(c419[k], minv[k]);
                if (bn < 0 || num < bnum) { bn = k; bnum = num; }
            }
        }
        int best;
        if (bn < 0) best = bs;
        else if (bs < 0) best = bn;
        else {
            const double fn = __ddiv_rn(bnum, V);
            best = fn < bsf ? bn : (bsf < fn ? bs : min(bn, bs));
        }
        ord[l] = (unsigned char)best;
        used |= 1u << best;
        last = best;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(ordbuf + ((long long)m * T.n_templates + t) * kMaxN);
    const uint4 *src = reinterpret_cast<const uint4 *>(ord);
    dst[0] = src[0]; dst[1] = src[1];
}

// ---- K: pairs --------------------------------------------------------------------
// Thread per (structure, template).  A pair is live iff every candidate set is
// non-empty (eff:Scanner.c:145-149).  Scan order:
//   reorder 0  template order
//   reorder 1  the reference's argsort(|C_k|, k) (eff:Scanner.c:88-99)
//   reorder 2  the engine's own (k_order above), read from ordbuf.
// The hit SET does not depend on the order (all constraints are pairwise and symmetric).
__global__ void __launch_bounds__(256) k_pairs(int n_mol, TemplateTable T, int n_specs, const CandHead *heads,
                                               const long long *atom_off, int reorder, const unsigned char *ordbuf,
                                               PairRec *recs, Counters *ctr)
{
    const long long total = (long long)n_mol * T.n_templates;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = false;
    int m = 0, t = 0, ta0 = 0, n = 0;
    int cnts[kMaxN];
    alignas(16) unsigned char ord[kMaxN];
    u64 sum = 0;
    if (p < total) {
        m = (int)(p / T.n_templates);
        t = (int)(p % T.n_templates);
        ta0 = T.atom_off[t];
        n = T.atom_off[t + 1] - ta0;
        live = n >= 1;
        for (int k = 0; k < n; k++) {
            const int c = heads[(long long)m * n_specs + T.spec[ta0 + k]].cnt;
            cnts[k] = c;
            live = live && (c > 0);
            sum += (u64)c;
        }
        if (live) {
            for (int k = 0; k < n; k++) ord[k] = (unsigned char)k;
            if (reorder == 1) {
                for (int i = 1; i < n; i++) {                   // stable insertion sort by (count, index)
                    const unsigned char v = ord[i];
                    const int cv = cnts[v];
                    int j = i - 1;
                    while (j >= 0 && cnts[ord[j]] > cv) { ord[j + 1] = ord[j]; j--; }
                    ord[j + 1] = v;
                }
            } else if (reorder == 2) {
                const uint4 *src = reinterpret_cast<const uint4 *>(ordbuf + p * kMaxN);
                uint4 *dst = reinterpret_cast<uint4 *>(ord);
                dst[0] = src[0]; dst[1] = src[1];
            }
        }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31;
    u64 wsum = live ? sum : 0;
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_down_sync(0xffffffffu, wsum, o);
    u64 base = 0;
    if (lane == 0 && mask) {
        base = atomicAdd(&ctr->live_pairs, (u64)__popc(mask));
        atomicAdd(&ctr->cand_atoms, wsum);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) {
        PairRec *r = recs + base + __popc(mask & ((1u << lane) - 1));
        // the 128-byte record leaves as eight 16-byte stores
        alignas(16) unsigned short spc[kMaxN];
        for (int k = 0; k < kMaxN; k++) {
            if (k >= n) ord[k] = 0;
            spc[k] = k < n ? (unsigned short)T.spec[ta0 + ord[k]] : (unsigned short)0;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(r);
        const long long ab = atom_off[m], sh0 = T.dist_off[t];
        dst[0] = make_uint4((unsigned)m, (unsigned)t, (unsigned)ta0, (unsigned)n);
        dst[1] = make_uint4((unsigned)(u64)ab, (unsigned)((u64)ab >> 32), (unsigned)(u64)sh0, (unsigned)((u64)sh0 >> 32));
        const uint4 *sp = reinterpret_cast<const uint4 *>(spc);
        const uint4 *op = reinterpret_cast<const uint4 *>(ord);
        dst[2] = sp[0]; dst[3] = sp[1]; dst[4] = sp[2]; dst[5] = sp[3];
        dst[6] = op[0]; dst[7] = op[1];
    }
}

// ---- K3: frontier expansion ---------------------------------------------------------
struct ExpandArgs {
    int level;                   // l = atoms already placed in the input items
    int n_specs;
    int tuple_stride;            // ints per tuple record: 1 + max template size
    int ignore_chain;
    const PairRec *recs;
    const ShellRec *shell;
    const AtomRec *atoms;
    const MolExt *ext;
    const CandHead *head;
    const unsigned short *xtab;  // x-bucket index of the lists (k_xindex)
    const float4 *cand;
    long long cand_cap;          // records allocated: lists are clamped to it while a chunk overflows (then re-run)
    const int *in;               // items [lp, a_0..a_{l-1}], stride l+1 (unused for l==0)
    const u64 *in_count;         // items available (may exceed in_cap after an overflow)
    long long in_cap;
    int *out;                    // items stride l+2
    u64 *out_count;
    long long out_cap;           // items
    int *tuples;
    long long tuple_cap;         // records
    Counters *ctr;
};

// TessTemplate_check for the new atom against one earlier atom
// (eff:TessTemplate.c:78-119): c/d chain differences (structure/template),
// a/b residue-number differences.
__device__ __forceinline__ bool pair_consistent(int ignore, bool mol_same_chain, bool tpl_same_chain, bool mol_same_res,
                                                bool tpl_same_res)
{
    if (ignore == 0) {
        if (mol_same_chain != tpl_same_chain) return false;
        if (!mol_same_chain) return true;
    }
    if (mol_same_res != tpl_same_res) return false;
    if (ignore != 2 && mol_same_res && !mol_same_chain) return false;
    return true;
}

struct alignas(16) ChainRec {    // one already-placed atom of a partial match, staged in shared memory
    double x, y, z;              // FP64 centre of the annulus
    double lo2, hi2;             // squared radii of the shell towards the atom being placed
    int res;                     // residue number of the placed atom
    int flag;                    // chain id | template flags << 8
};

struct alignas(16) GroupMeta {   // per partial match of the current batch
    long long abase;
    int lp;
    int complete;                // this level finishes the template
};

constexpr int kQueue = 96;       // per-warp survivor queue (entries): < 32 carried + 64 new

__host__ __device__ constexpr size_t expand_warp_bytes(int group, int level)
{
    const size_t ipw = 32 / group, lpad = level > 0 ? level : 1;
    return ipw * lpad * sizeof(ChainRec) + ipw * sizeof(GroupMeta) + (size_t)kQueue * sizeof(int2) +
           ((ipw * lpad + (size_t)kStageItems * (level + 2)) * sizeof(int) + 15) / 16 * 16;
}

// G lanes cooperate on one partial match; a warp handles 32/G of them per batch.
//  1. the lanes of a group stage the already-placed atoms (one 32-byte AtomRec and one
//     32-byte ShellRec each) into shared memory and reduce, with group shuffles, the
//     axis-aligned box of the annulus join (the reference's Join_computeBox,
//     eff:Join.c:82-111) in FP32, rounded OUTWARD by the record error bound;
//  2. the candidate list is x-sorted: a G-ary lower-bound search finds the first record
//     inside the box, the scan stops at the first record beyond it (this pair of steps
//     is what the k-d tree walk of eff:KdTree.c:238-296 becomes);
//  3. scan: one LDG.128 + six FP32 compares per candidate; box survivors are pushed to a
//     per-warp queue in shared memory (ballot + prefix popcount);
//  4. drain: ONE LANE PER SURVIVOR takes the exact FP64 predicate (_Annulus_po,
//     eff:Annulus.h:54-68, explicit RN ops, no FMA) against the staged atoms and the
//     chain/residue rules -- dense lanes instead of a divergent tail in the scan loop;
//  5. accepted extensions are compacted into a per-warp staging buffer that is flushed
//     to the next frontier with ONE global atomicAdd; finished tuples go to the tuple list.
template <int G, bool STATS>
__global__ void __launch_bounds__(kExpandThreads, 4) k_expand(const ExpandArgs A)
{
    constexpr int IPW = 32 / G;
    constexpr unsigned GM = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int l = A.level;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const int in_stride = l + 1, out_stride = l + 2;
    const int lpad = l > 0 ? l : 1;
    const unsigned lt = (1u << lane) - 1;

    unsigned char *wbase = smem_raw + (size_t)warp * expand_warp_bytes(G, l);
    ChainRec *chain = reinterpret_cast<ChainRec *>(wbase);
    GroupMeta *meta = reinterpret_cast<GroupMeta *>(chain + IPW * lpad);
    int2 *queue = reinterpret_cast<int2 *>(meta + IPW);
    int *chain_atom = reinterpret_cast<int *>(queue + kQueue);
    int *stage = chain_atom + IPW * lpad;
    int fill = 0, qfill = 0;

    if (A.ctr->tail_level != 0 && A.ctr->tail_level <= (u64)l) return;   // finished depth first (k_expand_tail)
    u64 n_in = *A.in_count;
    if (l > 0 && (long long)n_in > A.in_cap) n_in = (u64)A.in_cap;
    const long long total_warps = (long long)gridDim.x * kExpandWarps;
    u64 n_cand = 0, n_scan = 0, n_box = 0;           // lane 0 only
    const float INF = __int_as_float(0x7f800000);

    // exact predicate + emission for queue entries [q0, q0 + nq)
    auto drain = [&](int q0, int nq) {
        __syncwarp();
        bool ok = lane < nq;
        int a = 0, g = 0;
        GroupMeta gm; gm.abase = 0; gm.lp = 0; gm.complete = 0;
        if (ok) {
            const int2 e = queue[q0 + lane];
            a = e.x; g = e.y;
            gm = meta[g];
        }
        __syncwarp();   // the scan loop reuses these queue slots right after the drain (racecheck: WAR)
        if (ok && l > 0) {
            const AtomRec p = A.atoms[gm.abase + a];
            const ChainRec *ch = chain + g * lpad;
            // _Annulus_po (eff:Annulus.h:54-68): ((dx^2 + dy^2) + dz^2), centre - point
            for (int j = 0; j < l; j++) {
                const double t0 = __dsub_rn(ch[j].x, p.x), t1 = __dsub_rn(ch[j].y, p.y), t2 = __dsub_rn(ch[j].z, p.z);
                const double s = __dadd_rn(__dadd_rn(__dmul_rn(t0, t0), __dmul_rn(t1, t1)), __dmul_rn(t2, t2));
                if (s < ch[j].lo2 || s > ch[j].hi2) { ok = false; break; }
            }
            if (ok) {
                for (int j = 0; j < l; j++) {
                    const int f = ch[j].flag;
                    if (!pair_consistent(A.ignore_chain, (f & 0xff) == p.chain, (f & 0x100) != 0, ch[j].res == p.resseq,
                                         (f & 0x200) != 0)) { ok = false; break; }
                }
            }
        }
        const unsigned okm = __ballot_sync(FULL, ok);
        if (okm == 0) return;
        const unsigned cm = __ballot_sync(FULL, ok && gm.complete);
        const unsigned nm = okm & ~cm;
        if (cm) {   // finished tuples go straight to the tuple list (one atomic per warp)
            u64 tb = 0;
            if (lane == 0) {
                tb = atomicAdd(&A.ctr->tuple_slots, (u64)__popc(cm));
                atomicAdd(&A.ctr->tuples, (u64)__popc(cm));
                atomicAdd(&A.ctr->tuple_ints, (u64)__popc(cm) * (u64)(l + 1));
            }
            tb = __shfl_sync(FULL, tb, 0);
            if (ok && gm.complete) {
                const long long slot = (long long)tb + __popc(cm & lt);
                if (slot < A.tuple_cap) {
                    int *dst = A.tuples + slot * A.tuple_stride;
                    dst[0] = gm.lp;
                    for (int j = 0; j < l; j++) dst[1 + j] = chain_atom[g * lpad + j];
                    dst[1 + l] = a;
                }
            }
        }
        if (nm) {
            if (ok && !gm.complete) {
                int *dst = stage + (size_t)(fill + __popc(nm & lt)) * out_stride;
                dst[0] = gm.lp;
                for (int j = 0; j < l; j++) dst[1 + j] = chain_atom[g * lpad + j];
                dst[1 + l] = a;
            }
            fill += __popc(nm);
            if (fill + 32 > kStageItems) {   // flush
                __syncwarp();
                u64 ob = 0;
                if (lane == 0) ob = atomicAdd(A.out_count, (u64)fill);
                ob = __shfl_sync(FULL, ob, 0);
                if ((long long)ob < A.out_cap) {   // on overflow write the part that fits: every slot below
                    const int fit = (int)min((long long)fill, A.out_cap - (long long)ob);   // the capacity holds a real item
                    int *dst = A.out + (long long)ob * out_stride;
                    for (int i = lane; i < fit * out_stride; i += 32) dst[i] = stage[i];
                }
                fill = 0;
                __syncwarp();
            }
        }
    };

    for (long long base_item = ((long long)blockIdx.x * kExpandWarps + warp) * IPW; base_item < (long long)n_in;
         base_item += total_warps * IPW) {
        const long long item = base_item + grp;
        const bool valid = item < (long long)n_in;
        int cnt = 0;
        long long coff = 0;
        float bx0 = -INF, by0 = -INF, bz0 = -INF, bx1 = INF, by1 = INF, bz1 = INF;
        if (valid) {
            const int lp = l == 0 ? (int)item : A.in[item * in_stride];
            const PairRec *rec = A.recs + lp;
            const int4 h0 = *reinterpret_cast<const int4 *>(rec);            // m, t, ta0, n
            const longlong2 h1 = *reinterpret_cast<const longlong2 *>(&rec->abase);
            const int n = h0.w;
            const int k = rec->order[l];
            const CandHead hd = A.head[(long long)h0.x * A.n_specs + rec->spec[l]];
            cnt = hd.cnt;
            coff = hd.off;
            if (coff + cnt > A.cand_cap) cnt = (int)max(0ll, A.cand_cap - coff);   // overflowing chunk: stay in bounds
            if (gl == 0) {
                GroupMeta gm; gm.abase = h1.x; gm.lp = lp; gm.complete = (l + 1 == n) ? 1 : 0;
                meta[grp] = gm;
            }
            if (l > 0) {
                const MolExt me = A.ext[h0.x];
                const float err = __double2float_ru(me.ext * 2.4e-7 + 1e-30);  // > ext * 2^-22: record rounding
                for (int j = gl; j < l; j += G) {
                    const int a = A.in[item * in_stride + 1 + j];
                    const AtomRec ar = A.atoms[h1.x + a];
                    const ShellRec sh = A.shell[h1.y + (long long)rec->order[j] * n + k];
                    ChainRec c;
                    c.x = ar.x; c.y = ar.y; c.z = ar.z; c.lo2 = sh.lo2; c.hi2 = sh.hi2;
                    c.res = ar.resseq; c.flag = ar.chain | (sh.flags << 8);
                    chain[grp * lpad + j] = c;
                    chain_atom[grp * lpad + j] = a;
                    const double dx = __dsub_rn(ar.x, me.ox), dy = __dsub_rn(ar.y, me.oy), dz = __dsub_rn(ar.z, me.oz);
                    const float rb = sh.rb;
                    bx0 = fmaxf(bx0, __fsub_rd(__fsub_rd(__double2float_rd(dx), rb), err));
                    by0 = fmaxf(by0, __fsub_rd(__fsub_rd(__double2float_rd(dy), rb), err));
                    bz0 = fmaxf(bz0, __fsub_rd(__fsub_rd(__double2float_rd(dz), rb), err));
                    bx1 = fminf(bx1, __fadd_ru(__fadd_ru(__double2float_ru(dx), rb), err));
                    by1 = fminf(by1, __fadd_ru(__fadd_ru(__double2float_ru(dy), rb), err));
                    bz1 = fminf(bz1, __fadd_ru(__fadd_ru(__double2float_ru(dz), rb), err));
                }
            }
        }
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) {   // join box = intersection over the placed atoms
            bx0 = fmaxf(bx0, __shfl_xor_sync(FULL, bx0, o)); bx1 = fminf(bx1, __shfl_xor_sync(FULL, bx1, o));
            by0 = fmaxf(by0, __shfl_xor_sync(FULL, by0, o)); by1 = fminf(by1, __shfl_xor_sync(FULL, by1, o));
            bz0 = fmaxf(bz0, __shfl_xor_sync(FULL, bz0, o)); bz1 = fminf(bz1, __shfl_xor_sync(FULL, bz1, o));
        }
        // an empty join box kills the partial match without touching its candidates
        const bool boxed = valid && bx0 <= bx1 && by0 <= by1 && bz0 <= bz1;
        // ---- lower bound of bx0 in the x-sorted list: G-ary search, warp-uniform trip count
        int s_lo = 0, s_hi = cnt;
        const bool narrow = boxed && l > 0 && cnt > 2 * G;
        while (__any_sync(FULL, narrow && s_hi - s_lo > G)) {
            const bool act = narrow && s_hi - s_lo > G;
            const long long span = s_hi - s_lo;
            float x = 0.f;
            if (act) x = __ldg(&A.cand[coff + s_lo + (int)((span * (gl + 1)) / (G + 1))].x);
            const unsigned less = (__ballot_sync(FULL, act && x < bx0) >> (grp * G)) & GM;
            if (act) {
                const int t = __popc(less);               // probes are monotone: lanes 0..t-1 are below the box
                const int p_prev = s_lo + (int)((span * t) / (G + 1));
                const int p_next = s_lo + (int)((span * (t + 1)) / (G + 1));
                if (t < G) s_hi = p_next;
                if (t > 0) s_lo = p_prev + 1;
            }
        }
        // ---- scan ------------------------------------------------------------------------
        {   // algorithm-independent work counter: every candidate of the level counts once
            unsigned c_all = (valid && l > 0 && gl == 0) ? (unsigned)cnt : 0u;
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) c_all += __shfl_down_sync(FULL, c_all, o);
            if (lane == 0) n_cand += c_all;
        }
        bool live = boxed && s_lo < cnt;
        int c = s_lo + gl;
        while (__any_sync(FULL, live)) {
            // two candidates per lane and iteration: c and c + G (both loads in flight together)
            const bool in0 = live && c < cnt, in1 = live && c + G < cnt;
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
            if (in0) r0 = __ldg(A.cand + coff + c);
            if (in1) r1 = __ldg(A.cand + coff + c + G);
            const bool ok0 = in0 && r0.x >= bx0 && r0.x <= bx1 && r0.y >= by0 && r0.y <= by1 && r0.z >= bz0 && r0.z <= bz1;
            const bool ok1 = in1 && r1.x >= bx0 && r1.x <= bx1 && r1.y >= by0 && r1.y <= by1 && r1.z >= bz0 && r1.z <= bz1;
            const unsigned pm = __ballot_sync(FULL, (in0 && r0.x > bx1) || (in1 && r1.x > bx1));
            const unsigned m0 = __ballot_sync(FULL, ok0);
            const unsigned m1 = __ballot_sync(FULL, ok1);
            if (STATS && lane == 0) {
                n_scan += __popc(__ballot_sync(FULL, in0)) + __popc(__ballot_sync(FULL, in1));
                n_box += __popc(m0) + __popc(m1);
            } else if (STATS) { __ballot_sync(FULL, in0); __ballot_sync(FULL, in1); }
            if (m0 | m1) {
                const int k0 = __popc(m0);
                if (ok0) queue[qfill + __popc(m0 & lt)] = make_int2(__float_as_int(r0.w), grp);
                if (ok1) queue[qfill + k0 + __popc(m1 & lt)] = make_int2(__float_as_int(r1.w), grp);
                qfill += k0 + __popc(m1);
                if (qfill >= 64) { qfill -= 32; drain(qfill, 32); }
                if (qfill >= 32) { qfill -= 32; drain(qfill, 32); }
            }
            c += 2 * G;
            // sorted by x: nothing beyond the first record past the box can be inside
            if (((pm >> (grp * G)) & GM) || c - gl >= cnt) live = false;
        }
        if (qfill > 0) { drain(0, qfill); qfill = 0; }
        __syncwarp();   // the next batch overwrites chain/meta
    }
    __syncwarp();
    if (fill > 0) {
        u64 ob = 0;
        if (lane == 0) ob = atomicAdd(A.out_count, (u64)fill);
        ob = __shfl_sync(FULL, ob, 0);
        if ((long long)ob < A.out_cap) {
            const int fit = (int)min((long long)fill, A.out_cap - (long long)ob);
            int *dst = A.out + (long long)ob * out_stride;
            for (int i = lane; i < fit * out_stride; i += 32) dst[i] = stage[i];
        }
    }
    if (lane == 0) {   // statistics: lane 0 carries the warp totals
        if (n_cand) atomicAdd(&A.ctr->cand_tests, n_cand);
        if (n_scan) atomicAdd(&A.ctr->scan_tests, n_scan);
        if (n_box) atomicAdd(&A.ctr->box_pass, n_box);
    }
}

// ---- K4: superposition ----------------------------------------------------------------
struct HitBuf {
    int *pair;                   // PairRec index of the hit (reference-order mode), may be NULL
    int *mol, *tpl, *atoms;      // atoms stride = hit_stride
    double *rmsd, *rot, *c0, *c1;
    long long cap;
    int stride;
};

__device__ __forceinline__ void plane_rotate(double *W, double *P, int p, int q)
{
    // vendor/jess/src/Super.c:252-344: angle by atan2, then explicit update
    const double t = atan2(2.0 * W[3 * p + q], W[3 * q + q] - W[3 * p + p]) / 2.0;
    const double c = cos(t), s = sin(t);
    double wp[3], wq[3], pp[3], pq[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        wp[k] = c * W[3 * k + p] - s * W[3 * k + q];
        wq[k] = c * W[3 * k + q] + s * W[3 * k + p];
        pp[k] = c * P[3 * k + p] - s * P[3 * k + q];
        pq[k] = s * P[3 * k + p] + c * P[3 * k + q];
    }
    const double dpp = c * c * W[3 * p + p] + s * s * W[3 * q + q] - 2.0 * s * c * W[3 * p + q];
    const double dqq = s * s * W[3 * p + p] + c * c * W[3 * q + q] + 2.0 * s * c * W[3 * p + q];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        W[3 * k + p] = wp[k]; W[3 * p + k] = wp[k];
        W[3 * k + q] = wq[k]; W[3 * q + k] = wq[k];
        P[3 * k + p] = pp[k]; P[3 * k + q] = pq[k];
    }
    W[3 * p + p] = dpp; W[3 * q + q] = dqq;
    W[3 * p + q] = 0.0; W[3 * q + p] = 0.0;
}

// Thread per full tuple (grid-stride: the tuple count only exists on the device).
// Follows Superposition_compute/superpose (vendor/jess/src/Super.c:139-220, 417-538)
// including its summation order: the reference's pair list is built by prepending
// (Super.c:109-118), so every sum runs over template atoms n-1 .. 0.  Set 0 =
// structure atoms, set 1 = template positions (eff:Jess.c:189-192).
__global__ void __launch_bounds__(128) k_superpose(const int *tuples, const u64 *tuple_count, long long tuple_cap,
                                                   int tuple_stride, const PairRec *recs, TemplateTable T, MolTable M,
                                                   int mol_base, double rmsd_threshold, int all_tuples, HitBuf H,
                                                   Counters *ctr)
{
    u64 n_t = *tuple_count;
    if ((long long)n_t > tuple_cap) n_t = (u64)tuple_cap;
    const int lane = threadIdx.x & 31;
    const long long stride_all = (long long)gridDim.x * blockDim.x;
    for (long long wbase = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); wbase < (long long)n_t;
         wbase += stride_all) {
    const long long i = wbase + lane;
    bool hit = false, nan = false;
    double rmsd = 0.0, R[9], ca[3], cb[3];
    const int *tp = nullptr;
    const PairRec *rec = nullptr;
    int n = 0;
    unsigned char inv[kMaxN];   // template atom -> scan level
    if (i < (long long)n_t && tuples[i * tuple_stride] >= 0) {      // (a negative pair: a record its chunk left unused)
        tp = tuples + i * tuple_stride;
        rec = recs + tp[0];
        n = rec->n;
        for (int lv = 0; lv < n; lv++) inv[rec->order[lv]] = (unsigned char)lv;
        if (n == 1) {
            // Superposition_rmsd returns 0 for a single pair and never computes the rotation or
            // the centroids, which stay zeroed (vendor/jess/src/Super.c:125-130,152)
            for (int j = 0; j < 9; j++) R[j] = 0.0;
            for (int j = 0; j < 3; j++) ca[j] = cb[j] = 0.0;
            hit = all_tuples || 0.0 <= rmsd_threshold;
        } else {
        const long long abase = rec->abase;
        const double *tpos = T.pos + 3ll * rec->ta0;
        for (int j = 0; j < 3; j++) ca[j] = cb[j] = 0.0;
        for (int k = n - 1; k >= 0; k--) {
            const AtomRec ar = M.atoms[abase + tp[1 + inv[k]]];
            ca[0] += ar.x; ca[1] += ar.y; ca[2] += ar.z;
            cb[0] += tpos[3 * k]; cb[1] += tpos[3 * k + 1]; cb[2] += tpos[3 * k + 2];
        }
        for (int j = 0; j < 3; j++) { ca[j] /= (double)n; cb[j] /= (double)n; }
        double sa = 0.0, sb = 0.0, X[9];
        for (int j = 0; j < 9; j++) X[j] = 0.0;
        for (int k = n - 1; k >= 0; k--) {
            const AtomRec ar = M.atoms[abase + tp[1 + inv[k]]];
            const double a[3] = { ar.x - ca[0], ar.y - ca[1], ar.z - ca[2] };
            const double b[3] = { tpos[3 * k] - cb[0], tpos[3 * k + 1] - cb[1], tpos[3 * k + 2] - cb[2] };
#pragma unroll
            for (int j = 0; j < 3; j++) {
                sa += a[j] * a[j];
                sb += b[j] * b[j];
#pragma unroll
                for (int q = 0; q < 3; q++) X[3 * j + q] += a[j] * b[q];
            }
        }
        const double det = X[0] * (X[4] * X[8] - X[5] * X[7]) - X[1] * (X[3] * X[8] - X[5] * X[6]) +
                           X[2] * (X[3] * X[7] - X[4] * X[6]);
        double XX[9], P[9], Tm[9], e[3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < 3; q++) v += X[3 * q + a] * X[3 * q + b];
                XX[3 * a + b] = v;
            }
        for (int j = 0; j < 9; j++) P[j] = 0.0;
        P[0] = P[4] = P[8] = 1.0;
        for (int it = 0; it < 64; it++) {   // Super.c:346-414 (the reference has no sweep limit)
            double off = 0.0;
            off += fabs(XX[1]); off += fabs(XX[2]); off += fabs(XX[5]);
            if (off < 1e-12) break;
            plane_rotate(XX, P, 0, 1);
            plane_rotate(XX, P, 0, 2);
            plane_rotate(XX, P, 1, 2);
        }
        e[0] = fmax(XX[0], 0.0); e[1] = fmax(XX[4], 0.0); e[2] = fmax(XX[8], 0.0);
        double se = sqrt(e[0]) + sqrt(e[1]) + sqrt(e[2]);
        bool flip = false;
        if (det < 1e-8) { se -= 2.0 * sqrt(fmin(e[0], fmin(e[1], e[2]))); flip = true; }
        rmsd = sa + sb - 2.0 * se;
        rmsd = sqrt(fmax(rmsd, 0.0) / (double)n);
        if (all_tuples || rmsd <= rmsd_threshold) {
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    double f = sqrt(e[b]);
                    if (flip && b == 0) f = -f;
                    double v = 0.0;
#pragma unroll
                    for (int q = 0; q < 3; q++) v += X[3 * a + q] * P[3 * q + b] / f;
                    Tm[3 * a + b] = v;
                }
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    double v = 0.0;
#pragma unroll
                    for (int q = 0; q < 3; q++) v += P[3 * a + q] * Tm[3 * b + q];
                    R[3 * a + b] = v;
                    nan = nan || (v != v);
                }
            hit = !nan;
            if (all_tuples) { hit = true; nan = false; }   // the host layer applies Query.__next__'s filters
        }
        }   // n > 1
    }
    const unsigned hm = __ballot_sync(0xffffffffu, hit);
    const unsigned nm = __ballot_sync(0xffffffffu, nan);
    u64 hb = 0;
    if (lane == 0) {
        if (hm) hb = atomicAdd(&ctr->hits, (u64)__popc(hm));
        if (nm) atomicAdd(&ctr->nan_dropped, (u64)__popc(nm));
    }
    hb = __shfl_sync(0xffffffffu, hb, 0);
    if (hit) {
        const long long slot = (long long)hb + __popc(hm & ((1u << lane) - 1));
        if (slot < H.cap) {
            if (H.pair) H.pair[slot] = tp[0];
            H.mol[slot] = mol_base + rec->m;
            H.tpl[slot] = rec->t;
            for (int k = 0; k < H.stride; k++) H.atoms[slot * H.stride + k] = k < n ? tp[1 + inv[k]] : -1;
            H.rmsd[slot] = rmsd;
            for (int j = 0; j < 9; j++) H.rot[slot * 9 + j] = R[j];
            for (int j = 0; j < 3; j++) { H.c0[slot * 3 + j] = ca[j]; H.c1[slot * 3 + j] = cb[j]; }
        }
    }
    }   // grid-stride loop
}

}  // namespace jb
