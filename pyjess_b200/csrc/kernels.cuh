// kernels.cuh -- sm_100a device code of the template-query engine.
//
// One kernel per stage of the reference hot path:
//   k_classify        atoms x atom-specs -> per-(structure, spec) candidate
//                     lists (replaces TessTemplate_candidates, ResIndex_*,
//                     TessAtom_match: eff:TessTemplate.c:141-201,
//                     eff:ResIndex.c:28-153, eff:TessAtom.c:238-449)
//   k_shells          squared annulus radii per ordered template-atom pair
//                     (eff:Scanner.c:275-290, eff:Annulus.c:40-67)
//   k_pairs           per (structure, template): liveness + scan order
//                     (Scanner_reuse, eff:Scanner.c:106-183)
//   k_expand<G>       level-synchronous frontier expansion, G lanes per
//                     partial match (replaces Scanner_next + KdTreeQuery_next
//                     + _Join_ipo/_Annulus_po + TessTemplate_check:
//                     eff:Scanner.c:193-304, eff:KdTree.c:238-296,
//                     eff:Annulus.h:54-68, eff:TessTemplate.c:78-119)
//   k_superpose       batched FP64 superposition + RMSD filter + compaction
//                     (vendor/jess/src/Super.c:139-538, eff:Jess.c:176-195,
//                     src/pyjess/_jess.pyx:2238-2257)
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
};

struct alignas(32) AtomRec {   // one 32-byte sector per gathered atom
    double x, y, z;
    int resseq;
    int chain;
};

struct alignas(32) MolExt {    // frame of the FP32 candidate records of one structure
    double ox, oy, oz;         // origin (first atom)
    double ext;                // max |coordinate - origin| over the structure
};

struct MolTable {
    int n_molecules;           // of the chunk
    const long long *atom_off; // [M+1] (chunk view, absolute atom offsets)
    const AtomRec *atoms;      // packed at upload by k_pack_atoms
    const uint32_t *name, *res;
    MolExt *ext;               // [M] chunk-local, written by k_classify
};

struct alignas(16) CandHead {  // one candidate list = one (structure, spec)
    long long off;
    int cnt;
    int pad;
};

struct CandTable {
    CandHead *head;            // [M*S]
    float4 *rec;               // {x-ox, y-oy, z-oz, atom index bits}: FP32 is ONLY a conservative prefilter
    long long cap;             // capacity (records)
};

struct alignas(32) ShellRec {  // constraint of an ORDERED template-atom pair (earlier i, later k)
    double lo2, hi2;           // squared shell radii (eff:Scanner.c:275-290, eff:Annulus.c:40-67)
    int flags;                 // bit0: same template chain, bit1: same template residue number
    int pad;
};

struct alignas(16) PairRec {   // one live (structure, template) pair, 128 bytes
    int m;                     // structure index within the chunk
    int t;                     // template index
    int ta0;                   // first template atom
    int n;                     // template size
    long long abase;           // first atom of the structure
    long long shell0;          // first ShellRec of the template
    unsigned short spec[kMaxN];// spec id of the atom scanned at level l
    unsigned char order[kMaxN];// scan order (eff:Scanner.c:88-99,160-174)
};

struct Counters {              // device counters of one chunk attempt
    u64 cand_total;            // records reserved in CandTable
    u64 live_pairs;
    u64 cand_atoms;            // sum_k |C_k| over live pairs
    u64 cand_tests;            // (partial match, candidate) evaluations
    u64 box_pass;              // ... that survived the FP32 join-box prefilter
    u64 tuples;                // full tuples produced
    u64 tuple_ints;            // sum of their sizes
    u64 hits;                  // hits written (absolute, across chunks)
    u64 nan_dropped;
    u64 level[kMaxN + 1];      // level[l] = items with l atoms (l>=1)
    u64 unsupported;           // pairs skipped because n == 1
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

// ---- K1: candidate lists ------------------------------------------------------
// One CTA per structure.  Warp w owns a contiguous slice of the atoms; pass 1
// counts matches per (warp, spec) in shared memory, a scan over warps gives every
// warp its write offset, pass 2 re-evaluates and writes the records in ascending
// atom order (deterministic lists).  List placement in the record array is
// reserved with one atomicAdd per (structure, spec).  Pass 1 also measures the
// extent of the structure around its first atom: the FP32 records are stored
// relative to that origin and the extent bounds their rounding error.
__global__ void __launch_bounds__(256) k_classify(MolTable M, SpecTable S, CandTable C, Counters *ctr)
{
    extern __shared__ int sh_cnt[];                 // [warps][S]
    __shared__ double sh_ext[8];
    const int m = blockIdx.x;
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ns = S.n_specs;
    const long long a0 = M.atom_off[m], a1 = M.atom_off[m + 1];
    const int na = (int)(a1 - a0);
    double ox = 0.0, oy = 0.0, oz = 0.0, ext = 0.0;
    if (na > 0) { const AtomRec o = M.atoms[a0]; ox = o.x; oy = o.y; oz = o.z; }
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
                name = M.name[a0 + a];
                name_up = upper4(name);
                res_up = upper4(M.res[a0 + a]);
                const AtomRec ar = M.atoms[a0 + a];
                const double dx = __dsub_rn(ar.x, ox), dy = __dsub_rn(ar.y, oy), dz = __dsub_rn(ar.z, oz);
                if (pass == 0) ext = fmax(ext, fmax(fabs(dx), fmax(fabs(dy), fabs(dz))));
                rec = make_float4(__double2float_rn(dx), __double2float_rn(dy), __double2float_rn(dz), __int_as_float(a));
            }
            for (int s = 0; s < ns; s++) {
                const bool ok = valid && spec_match(S, s, name, name_up, res_up);
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                if (mask == 0) continue;
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
            }
        }
        if (pass == 0) {
            for (int o = 16; o > 0; o >>= 1) ext = fmax(ext, __shfl_xor_sync(0xffffffffu, ext, o));
            if (lane == 0) sh_ext[w] = ext;
            __syncthreads();
            if (threadIdx.x == 0) {
                double e = 0.0;
                for (int v = 0; v < nw; v++) e = fmax(e, sh_ext[v]);
                MolExt me; me.ox = ox; me.oy = oy; me.oz = oz; me.ext = e;
                M.ext[m] = me;
            }
            // exclusive scan over warps, reserve the list, publish count/offset
            for (int s = threadIdx.x; s < ns; s += blockDim.x) {
                int run = 0;
                for (int v = 0; v < nw; v++) {
                    const int c = sh_cnt[v * ns + s];
                    sh_cnt[v * ns + s] = run;
                    run += c;
                }
                CandHead h; h.off = 0; h.cnt = run; h.pad = 0;
                if (run > 0) h.off = (long long)atomicAdd(&ctr->cand_total, (u64)run);
                C.head[(long long)m * ns + s] = h;
            }
            __syncthreads();
        }
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
        r.pad = 0;
        T.shell[d0 + e] = r;
    }
}

// ---- K: pairs --------------------------------------------------------------------
// Thread per (structure, template).  A pair is live iff every candidate set is
// non-empty (eff:Scanner.c:145-149); the scan order is argsort(|C_k|, k)
// (eff:Scanner.c:88-99) or the identity when reorder is off.
__global__ void __launch_bounds__(256) k_pairs(int n_mol, TemplateTable T, int n_specs, const CandHead *heads,
                                               const long long *atom_off, int reorder, PairRec *recs, Counters *ctr)
{
    const long long total = (long long)n_mol * T.n_templates;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = false;
    int m = 0, t = 0, ta0 = 0, n = 0;
    int cnts[kMaxN];
    unsigned char ord[kMaxN];
    u64 sum = 0;
    bool single = false;
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
        if (live && n == 1) { single = true; live = false; }    // see DESIGN.md: n==1 not on the GPU path
        if (live) {
            for (int k = 0; k < n; k++) ord[k] = (unsigned char)k;
            if (reorder) {
                for (int i = 1; i < n; i++) {                   // stable insertion sort by (count, index)
                    const unsigned char v = ord[i];
                    const int cv = cnts[v];
                    int j = i - 1;
                    while (j >= 0 && cnts[ord[j]] > cv) { ord[j + 1] = ord[j]; j--; }
                    ord[j + 1] = v;
                }
            }
        }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31;
    u64 wsum = live ? sum : 0;
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_down_sync(0xffffffffu, wsum, o);
    const unsigned smask = __ballot_sync(0xffffffffu, single);
    u64 base = 0;
    if (lane == 0) {
        if (mask) {
            base = atomicAdd(&ctr->live_pairs, (u64)__popc(mask));
            atomicAdd(&ctr->cand_atoms, wsum);
        }
        if (smask) atomicAdd(&ctr->unsupported, (u64)__popc(smask));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) {
        PairRec *r = recs + base + __popc(mask & ((1u << lane) - 1));
        r->m = m; r->t = t; r->ta0 = ta0; r->n = n;
        r->abase = atom_off[m];
        r->shell0 = T.dist_off[t];
        for (int k = 0; k < kMaxN; k++) {
            r->order[k] = k < n ? ord[k] : 0;
            r->spec[k] = k < n ? (unsigned short)T.spec[ta0 + ord[k]] : 0;
        }
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
    const float4 *cand;
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

// G lanes cooperate on one partial match; a warp handles 32/G of them at a time.
//  1. the lanes of a group stage the already-placed atoms (one 32-byte AtomRec and
//     one 32-byte ShellRec each) into shared memory and reduce the axis-aligned box
//     of the annulus join (the reference's Join_computeBox, eff:Join.c:82-111) with
//     group shuffles, in FP32, rounded OUTWARD by the record error bound;
//  2. they stride the candidate list: one LDG.128 + six FP32 compares per candidate;
//  3. only box survivors take the exact FP64 predicate (_Annulus_po,
//     eff:Annulus.h:54-68, explicit RN ops, no FMA) and the chain/residue rules;
//  4. survivors are compacted (ballot + prefix popcount) into a per-warp staging
//     buffer that is flushed to the next frontier with ONE global atomicAdd.
template <int G>
__global__ void __launch_bounds__(kExpandThreads, 2) k_expand(const ExpandArgs A)
{
    constexpr int IPW = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int l = A.level;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const int in_stride = l + 1, out_stride = l + 2;
    const int lpad = l > 0 ? l : 1;

    // shared memory carve-up (per warp): doubles first (alignment), then ints
    const size_t dbl_per_warp = (size_t)IPW * lpad * 5;
    const size_t int_per_warp = (size_t)IPW * lpad * 3 + (size_t)kStageItems * out_stride;
    double *wd = reinterpret_cast<double *>(smem_raw) + warp * dbl_per_warp;
    int *wi = reinterpret_cast<int *>(smem_raw + sizeof(double) * dbl_per_warp * kExpandWarps) + warp * int_per_warp;
    double *sx = wd + (size_t)grp * lpad * 5, *sy = sx + lpad, *sz = sy + lpad, *slo = sz + lpad, *shi = slo + lpad;
    int *sa = wi + (size_t)grp * lpad * 3, *sres = sa + lpad, *sflag = sres + lpad;
    int *stage = wi + (size_t)IPW * lpad * 3;
    int fill = 0;

    u64 n_in = *A.in_count;
    if (l > 0 && (long long)n_in > A.in_cap) n_in = (u64)A.in_cap;
    const long long total_warps = (long long)gridDim.x * kExpandWarps;
    unsigned tests = 0, boxed = 0;
    const float INF = __int_as_float(0x7f800000);

    for (long long base_item = ((long long)blockIdx.x * kExpandWarps + warp) * IPW; base_item < (long long)n_in;
         base_item += total_warps * IPW) {
        const long long item = base_item + grp;
        const bool valid = item < (long long)n_in;
        int lp = 0, n = 0, k = 0, cnt = 0;
        long long coff = 0, abase = 0;
        float bx0 = -INF, by0 = -INF, bz0 = -INF, bx1 = INF, by1 = INF, bz1 = INF;
        __syncwarp();   // previous iteration's readers are done with the scratch
        if (valid) {
            lp = l == 0 ? (int)item : A.in[item * in_stride];
            const PairRec *rec = A.recs + lp;
            const int4 h0 = *reinterpret_cast<const int4 *>(rec);            // m, t, ta0, n
            const longlong2 h1 = *reinterpret_cast<const longlong2 *>(&rec->abase);
            n = h0.w;
            abase = h1.x;
            k = rec->order[l];
            const CandHead hd = A.head[(long long)h0.x * A.n_specs + rec->spec[l]];
            cnt = hd.cnt;
            coff = hd.off;
            if (l > 0) {
                const MolExt me = A.ext[h0.x];
                const double err = me.ext * 2.4e-7 + 1e-30;                  // > ext * 2^-22: record rounding
                for (int j = gl; j < l; j += G) {
                    const int a = A.in[item * in_stride + 1 + j];
                    const AtomRec ar = A.atoms[abase + a];
                    const ShellRec sh = A.shell[h1.y + (long long)rec->order[j] * n + k];
                    sa[j] = a;
                    sx[j] = ar.x; sy[j] = ar.y; sz[j] = ar.z;
                    slo[j] = sh.lo2; shi[j] = sh.hi2;
                    sres[j] = ar.resseq;
                    sflag[j] = ar.chain | (sh.flags << 8);
                    // outer radius, rounded up with a generous relative margin
                    const double b = __dmul_ru(__dsqrt_ru(sh.hi2), 1.000000001);
                    bx0 = fmaxf(bx0, __double2float_rd(__dsub_rd(__dsub_rd(__dsub_rd(ar.x, b), me.ox), err)));
                    by0 = fmaxf(by0, __double2float_rd(__dsub_rd(__dsub_rd(__dsub_rd(ar.y, b), me.oy), err)));
                    bz0 = fmaxf(bz0, __double2float_rd(__dsub_rd(__dsub_rd(__dsub_rd(ar.z, b), me.oz), err)));
                    bx1 = fminf(bx1, __double2float_ru(__dadd_ru(__dsub_ru(__dadd_ru(ar.x, b), me.ox), err)));
                    by1 = fminf(by1, __double2float_ru(__dadd_ru(__dsub_ru(__dadd_ru(ar.y, b), me.oy), err)));
                    bz1 = fminf(bz1, __double2float_ru(__dadd_ru(__dsub_ru(__dadd_ru(ar.z, b), me.oz), err)));
                }
            }
        }
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) {   // join box = intersection over the placed atoms
            bx0 = fmaxf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); bx1 = fminf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
            by0 = fmaxf(by0, __shfl_xor_sync(0xffffffffu, by0, o)); by1 = fminf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
            bz0 = fmaxf(bz0, __shfl_xor_sync(0xffffffffu, bz0, o)); bz1 = fminf(bz1, __shfl_xor_sync(0xffffffffu, bz1, o));
        }
        __syncwarp();
        const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
        const bool complete = (l + 1 == n);
        for (int c0 = 0; c0 < maxcnt; c0 += G) {
            const int c = c0 + gl;
            bool ok = valid && c < cnt;
            int a = 0;
            if (ok) {
                tests++;
                const float4 cr = __ldg(A.cand + coff + c);
                a = __float_as_int(cr.w);
                ok = cr.x >= bx0 && cr.x <= bx1 && cr.y >= by0 && cr.y <= by1 && cr.z >= bz0 && cr.z <= bz1;
            }
            if (ok && l > 0) {
                boxed++;
                const AtomRec p = A.atoms[abase + a];
                // _Annulus_po (eff:Annulus.h:54-68): ((dx^2 + dy^2) + dz^2), centre - point
                for (int j = 0; j < l; j++) {
                    const double t0 = __dsub_rn(sx[j], p.x), t1 = __dsub_rn(sy[j], p.y), t2 = __dsub_rn(sz[j], p.z);
                    const double s = __dadd_rn(__dadd_rn(__dmul_rn(t0, t0), __dmul_rn(t1, t1)), __dmul_rn(t2, t2));
                    if (s < slo[j] || s > shi[j]) { ok = false; break; }
                }
                if (ok) {
                    for (int j = 0; j < l; j++) {
                        const int f = sflag[j];
                        if (!pair_consistent(A.ignore_chain, (f & 0xff) == p.chain, (f & 0x100) != 0, sres[j] == p.resseq,
                                             (f & 0x200) != 0)) { ok = false; break; }
                    }
                }
            }
            const unsigned okm = __ballot_sync(0xffffffffu, ok);
            if (okm == 0) continue;
            const unsigned lt = (1u << lane) - 1;
            const unsigned cm = __ballot_sync(0xffffffffu, ok && complete);
            const unsigned nm = okm & ~cm;
            if (cm) {   // finished tuples go straight to the tuple list (one atomic per warp)
                u64 tb = 0;
                if (lane == 0) {
                    tb = atomicAdd(&A.ctr->tuples, (u64)__popc(cm));
                    atomicAdd(&A.ctr->tuple_ints, (u64)__popc(cm) * (u64)(l + 1));
                }
                tb = __shfl_sync(0xffffffffu, tb, 0);
                if (ok && complete) {
                    const long long slot = (long long)tb + __popc(cm & lt);
                    if (slot < A.tuple_cap) {
                        int *dst = A.tuples + slot * A.tuple_stride;
                        dst[0] = lp;
                        for (int j = 0; j < l; j++) dst[1 + j] = sa[j];
                        dst[1 + l] = a;
                    }
                }
            }
            if (nm) {
                if (ok && !complete) {
                    int *dst = stage + (size_t)(fill + __popc(nm & lt)) * out_stride;
                    dst[0] = lp;
                    for (int j = 0; j < l; j++) dst[1 + j] = sa[j];
                    dst[1 + l] = a;
                }
                fill += __popc(nm);
                if (fill + 32 > kStageItems) {   // flush
                    __syncwarp();
                    u64 ob = 0;
                    if (lane == 0) ob = atomicAdd(A.out_count, (u64)fill);
                    ob = __shfl_sync(0xffffffffu, ob, 0);
                    if ((long long)ob + fill <= A.out_cap) {
                        int *dst = A.out + (long long)ob * out_stride;
                        for (int i = lane; i < fill * out_stride; i += 32) dst[i] = stage[i];
                    }
                    fill = 0;
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    if (fill > 0) {
        u64 ob = 0;
        if (lane == 0) ob = atomicAdd(A.out_count, (u64)fill);
        ob = __shfl_sync(0xffffffffu, ob, 0);
        if ((long long)ob + fill <= A.out_cap) {
            int *dst = A.out + (long long)ob * out_stride;
            for (int i = lane; i < fill * out_stride; i += 32) dst[i] = stage[i];
        }
    }
    // statistics: one atomic per warp
    u64 t64 = tests, b64 = boxed;
    for (int o = 16; o > 0; o >>= 1) {
        t64 += __shfl_down_sync(0xffffffffu, t64, o);
        b64 += __shfl_down_sync(0xffffffffu, b64, o);
    }
    if (lane == 0 && t64) {
        if (l > 0) atomicAdd(&A.ctr->cand_tests, t64);
        atomicAdd(&A.ctr->box_pass, b64);
    }
}

// ---- K4: superposition ----------------------------------------------------------------
struct HitBuf {
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
                                                   int mol_base, double rmsd_threshold, HitBuf H, Counters *ctr)
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
    if (i < (long long)n_t) {
        tp = tuples + i * tuple_stride;
        rec = recs + tp[0];
        n = rec->n;
        for (int lv = 0; lv < n; lv++) inv[rec->order[lv]] = (unsigned char)lv;
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
        if (rmsd <= rmsd_threshold) {
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
        }
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
