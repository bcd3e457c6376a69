// prepare_small.cuh -- set-up kernels of the LATENCY path (small batches: a single-structure query).
//
// The batch kernels are organised for throughput: one CTA bitonic-sorts a structure's atoms, one
// THREAD scans the expected fan-out of a (structure, template) pair.  With one structure and a few
// templates those serial stretches ARE the run time (profiles/r2: k_sort_atoms 70 us, k_order 25 us
// of a 0.36 ms query).  Here the same results come from wider, shorter kernels:
//   k_prepare_small<ITEMS>  atom records + frame + x-sorted view of a structure by ONE block radix sort
//                           (replaces k_pack_atoms + k_sort_atoms; same order: the sort is stable and
//                           starts from atom order, the bitonic keys break x ties by atom index)
//   k_pairs_small           liveness + scan order + pair record by ONE WARP per pair, lane k = template
//                           atom k (replaces k_order + k_pairs; the same FP64 operations per atom,
//                           argmin with lowest-index ties == the serial "first strictly smaller" scan)
#pragma once
#include <cub/block/block_radix_sort.cuh>
#include "kernels.cuh"

namespace jb {

constexpr int kPrepThreads = 1024;

template <int ITEMS>
__global__ void __launch_bounds__(kPrepThreads) k_prepare_small(const long long *atom_off, AtomRec *atoms, const uint32_t *name,
                                                               const uint32_t *res, MolExt *ext, float4 *srec, uint32_t *sname,
                                                               uint32_t *sres, RawAtoms raw)
{
    using Sort = cub::BlockRadixSort<unsigned, kPrepThreads, ITEMS, unsigned>;
    __shared__ typename Sort::TempStorage tmp;
    __shared__ double sh_ext[kPrepThreads / 32];
    const int m = blockIdx.x;
    const long long a0 = atom_off[m], a1 = atom_off[m + 1];
    const int na = (int)(a1 - a0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double ox = 0.0, oy = 0.0, oz = 0.0, e = 0.0;
    if (na > 0) { ox = raw.x[a0]; oy = raw.y[a0]; oz = raw.z[a0]; }
    unsigned key[ITEMS], val[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int a = tid * ITEMS + i;
        key[i] = 0xffffffffu; val[i] = 0xffffffffu;
        if (a < na) {
            AtomRec r;
            r.x = raw.x[a0 + a]; r.y = raw.y[a0 + a]; r.z = raw.z[a0 + a];
            r.resseq = raw.resseq[a0 + a];
            r.chain = (int)raw.chain[a0 + a];
            atoms[a0 + a] = r;
            e = fmax(e, fmax(fabs(__dsub_rn(r.x, ox)), fmax(fabs(__dsub_rn(r.y, oy)), fabs(__dsub_rn(r.z, oz)))));
            key[i] = float_orderable(__double2float_rn(__dsub_rn(r.x, ox)));
            val[i] = (unsigned)a;
        }
    }
    for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
    if (lane == 0) sh_ext[w] = e;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int i = 0; i < kPrepThreads / 32; i++) v = fmax(v, sh_ext[i]);
        MolExt me; me.ox = ox; me.oy = oy; me.oz = oz; me.ext = v;
        ext[m] = me;
    }
    Sort(tmp).Sort(key, val);                      // stable: equal x keep atom order; padding (index >= na) stays last
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int pos = tid * ITEMS + i;
        if (pos < na) {
            const int a = (int)val[i];
            srec[a0 + pos] = make_float4(__double2float_rn(__dsub_rn(raw.x[a0 + a], ox)), __double2float_rn(__dsub_rn(raw.y[a0 + a], oy)),
                                         __double2float_rn(__dsub_rn(raw.z[a0 + a], oz)), __int_as_float(a));
            sname[a0 + pos] = name[a0 + a];
            sres[a0 + pos] = res[a0 + a];
        }
    }
}

// argmin over the lanes with `on`: smallest value, lowest lane on ties; returns the lane or -1
__device__ __forceinline__ int warp_argmin(double v, bool on, double *vmin)
{
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double best = on ? v : inf;
    int who = on ? (int)(threadIdx.x & 31) : 64;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int ow = __shfl_xor_sync(0xffffffffu, who, o);
        if (ow < 64 && (who == 64 || ob < best || (ob == best && ow < who))) { best = ob; who = ow; }
    }
    *vmin = best;
    return who == 64 ? -1 : who;
}

constexpr int kPairWarps = 4;

__global__ void __launch_bounds__(kPairWarps * 32) k_pairs_small(int n_mol, TemplateTable T, int n_specs, const CandHead *heads,
                                                                 const long long *atom_off, int reorder, PairRec *recs, Counters *ctr)
{
    __shared__ unsigned char s_ord[kPairWarps][kMaxN];
    __shared__ double s_vol[kPairWarps][kMaxN * kMaxN];      // the template's volume table and same-residue flags, staged
    __shared__ unsigned char s_same[kPairWarps][kMaxN * kMaxN];   // once: the steps below must not wait on global memory
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long total = (long long)n_mol * T.n_templates;
    const long long p = (long long)blockIdx.x * kPairWarps + wid;
    if (p >= total) return;
    const int m = (int)(p / T.n_templates), t = (int)(p % T.n_templates);
    const int ta0 = T.atom_off[t], n = T.atom_off[t + 1] - ta0;
    const long long d0 = T.dist_off[t];
    const bool mine = lane < n;
    const int cnt = mine ? heads[(long long)m * n_specs + T.spec[ta0 + lane]].cnt : 1;
    if (n < 1 || __any_sync(0xffffffffu, mine && cnt <= 0)) return;      // eff:Scanner.c:145-149: an empty candidate set kills the pair
    int myord = lane;                                                     // reorder 0: template order
    if (reorder == 1) {                                                   // eff:Scanner.c:88-99: stable argsort(|C_k|, k)
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const int cj = __shfl_sync(0xffffffffu, cnt, j);
            if (mine && (cj < cnt || (cj == cnt && j < lane))) rank++;
        }
        if (mine) s_ord[wid][rank] = (unsigned char)lane;
        __syncwarp();
        if (mine) myord = s_ord[wid][lane];
    } else if (reorder == 2) {                                            // the engine's expected fan-out order: k_order, lane k = atom k
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        const double V = __dmul_rn(19.0, (double)(atom_off[m + 1] - atom_off[m]));
        const double c419 = __dmul_rn((double)cnt, 4.19);
        for (int e = lane; e < n * n; e += 32) {
            s_vol[wid][e] = T.vol[d0 + e];
            s_same[wid][e] = (T.shell[d0 + e].flags & 3) == 3;
        }
        __syncwarp();
        // first atom: lane c scores itself
        double mnum = inf;
        for (int k = 0; k < n; k++) {
            const double ck = __shfl_sync(0xffffffffu, c419, k);
            if (mine && k != lane && !s_same[wid][lane * n + k]) {
                const double num = __dmul_rn(ck, s_vol[wid][lane * n + k]);
                if (num < mnum) mnum = num;
            }
        }
        const double mf = mnum == inf ? 1.0 : __ddiv_rn(mnum, V);
        const double sc = __dmul_rn((double)cnt, __dadd_rn(1.0, mf));
        double dummy;
        int last = warp_argmin(sc, mine, &dummy);
        bool used = lane == last, same = false;
        double minv = inf;
        if (lane == 0) myord = last;
        for (int l = 1; l < n; l++) {
            const bool open = mine && !used;
            if (open) {
                const double v = s_vol[wid][last * n + lane];
                if (v < minv) minv = v;
                if (s_same[wid][last * n + lane]) same = true;
            }
            double bsf, bnum;
            const int bs = warp_argmin(__dadd_rn(1e3, (double)cnt), open && same, &bsf);
            const int bn = warp_argmin(__dmul_rn(c419, minv), open && !same, &bnum);
            int best;
            if (bn < 0) best = bs;
            else if (bs < 0) best = bn;
            else {
                const double fn = __ddiv_rn(bnum, V);
                best = fn < bsf ? bn : (bsf < fn ? bs : min(bn, bs));
            }
            if (lane == l) myord = best;
            if (lane == best) used = true;
            last = best;
        }
    }
    if (!mine) myord = 0;
    // the record (layout of k_pairs)
    u64 sum = mine ? (u64)cnt : 0;
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    u64 slot = 0;
    if (lane == 0) {
        slot = atomicAdd(&ctr->live_pairs, 1ull);
        atomicAdd(&ctr->cand_atoms, sum);
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    PairRec *r = recs + slot;
    if (lane == 0) {
        r->m = m; r->t = t; r->ta0 = ta0; r->n = n;
        r->abase = atom_off[m];
        r->shell0 = d0;
    }
    r->spec[lane] = mine ? (unsigned short)T.spec[ta0 + myord] : (unsigned short)0;
    r->order[lane] = (unsigned char)myord;
}

}  // namespace jb
