// expand_dfs.cuh -- depth-first frontier sweep for SMALL batches (one launch for all levels).
//
// The level-synchronous sweep (expand_grid.cuh) costs one kernel launch per template atom: ~17
// launches for a cfg 3 template set, each with a few microseconds of launch + drain latency --
// nothing at 10^6 pairs per step, most of the run time of a single-structure query (the
// reference answers BASELINE config 1 in 0.2 ms; the README's thread-pool recipe issues exactly
// such queries).  k_expand_dfs walks the SAME search tree depth first, like the reference's
// Scanner_next (eff:Scanner.c:193-304) does.  A step of a depth-first walk is a chain of dependent
// loads (~1-2 us on a GPU), so the tree of a pair is cut at its ROOTS: work item = (live pair, slice
// s of kRootSlices) and the warp that takes it walks the subtrees of the first-atom candidates s,
// s + kRootSlices, ... -- a few partial matches each.  Below the root the 32 lanes test 32 records
// of the current level's candidate list at a time, accepted records wait in a per-level queue and
// are descended into one by one.  Every decision is taken exactly as in
// the level kernels (FP32 image with a rigorous error band, the reference's FP64 operations for
// the borderline records, TessTemplate_check): the tuple SET and the work counters are identical,
// only the order in which tuples reach the tuple list differs (it is re-ordered later anyway).
#pragma once
#include "kernels.cuh"

namespace jb {

constexpr int kDfsWarps = 4;
constexpr int kRootSlices = 32;              // work items per live pair
constexpr int kDfsThreads = kDfsWarps * 32;

struct alignas(16) DfsWarp {                 // per-warp shared memory
    float4 cf[kMaxN], th[kMaxN];             // constraints of the level being scanned: one entry per placed atom
    double clo2[kMaxN], chi2[kMaxN];
    int cres[kMaxN], cflag[kMaxN];
    int atom[kMaxN];                         // placed atoms, by level
    int queue[kMaxN][32];                    // accepted records (atom indices) of every level, not yet descended into
    int qn[kMaxN], qi[kMaxN];
    int next[kMaxN], end[kMaxN];             // record range of the level's list still to be scanned
    long long off[kMaxN];                    // first record of the level's list
    unsigned count[kMaxN + 2];               // accepted partial matches per size (work counters), [kMaxN + 1]: unused
};

// TAIL = false: the whole tree of every live pair (small chunks, above).
// TAIL = true (k_expand_tail): the level-synchronous sweep hands over its frontier of level l0 -- items
// [pair, a_0 .. a_{l0-1}] -- once that frontier is small: the last levels of a large batch hold a few thousand partial
// matches each (cfg 3: 4,655 at level 9 of 17) and cost ~40 us of launch + dependent-load latency per level whatever
// their size.  One warp per item finishes its subtree depth first; the level kernels of levels >= l0 then return at once
// (Counters::tail_level).  Every block takes the same decision from the same final counter; the flag only tells LATER
// kernels.
template <bool TAIL>
__global__ void __launch_bounds__(kDfsThreads) k_expand_dfs_t(const ExpandArgs A, const int l0, const long long tail_max)
{
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ DfsWarp sh[kDfsWarps];
    DfsWarp &W = sh[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    const long long total_warps = (long long)gridDim.x * kDfsWarps;
    u64 n_in = *A.in_count;                  // live pairs (TAIL: partial matches of l0 atoms)
    if (TAIL) {
        const u64 fired = A.ctr->tail_level;
        if ((fired != 0 && fired < (u64)l0) || (long long)n_in > tail_max || (long long)n_in > A.in_cap) return;
        if (blockIdx.x == 0 && threadIdx.x == 0) A.ctr->tail_level = (u64)l0;
    }
    if ((long long)n_in > A.in_cap) n_in = (u64)A.in_cap;
    const float INF = __int_as_float(0x7f800000);
    u64 n_cand = 0;
    for (int i = lane; i < kMaxN + 2; i += 32) W.count[i] = 0u;
    __syncwarp();

    for (long long item = (long long)blockIdx.x * kDfsWarps + (threadIdx.x >> 5); item < (long long)n_in * (TAIL ? 1 : kRootSlices);
         item += total_warps) {
        const long long lp = TAIL ? (long long)A.in[item * (l0 + 1)] : item / kRootSlices;
        const int slice = TAIL ? 0 : (int)(item % kRootSlices);
        const PairRec *rec = A.recs + lp;
        const int4 h0 = *reinterpret_cast<const int4 *>(rec);            // m, t, ta0, n
        const longlong2 h1 = *reinterpret_cast<const longlong2 *>(&rec->abase);
        const int n = h0.w;
        const MolExt me = A.ext[h0.x];
        const float err = __double2float_ru(me.ext * 2.4e-7 + 1e-30);    // > ext * 2^-22: record rounding
        const float extu = __double2float_ru(me.ext);
        int l = 0;                           // atoms placed
        int tab_lv = -1;                     // level whose constraint tables are in shared memory

        // (re)build the constraint tables of level `lv` from the placed atoms; returns the join box
        auto tables = [&](const int lv, float &bx0, float &bx1) {
            bx0 = -INF; bx1 = INF;
            float by0 = -INF, bz0 = -INF, by1 = INF, bz1 = INF;
            const int k = rec->order[lv];
            for (int j = lane; j < lv; j += 32) {
                const int a = W.atom[j];
                const AtomRec ar = A.atoms[h1.x + a];
                const ShellRec shl = A.shell[h1.y + (long long)rec->order[j] * n + k];
                W.clo2[j] = shl.lo2; W.chi2[j] = shl.hi2;
                W.cres[j] = ar.resseq; W.cflag[j] = ar.chain | (shl.flags << 8);
                const double dx = __dsub_rn(ar.x, me.ox), dy = __dsub_rn(ar.y, me.oy), dz = __dsub_rn(ar.z, me.oz);
                const float rb = shl.rb;
                const float eb = __fmul_ru(__fmul_ru(9.5367431640625e-07f, rb), __fadd_ru(extu, rb));   // 2^-20 rb (ext + rb)
                W.cf[j] = make_float4(__double2float_rn(dx), __double2float_rn(dy), __double2float_rn(dz),
                                      __fsub_rd(__double2float_rd(shl.lo2), eb));
                W.th[j] = make_float4(__fadd_ru(__double2float_ru(shl.lo2), eb), __fsub_rd(__double2float_rd(shl.hi2), eb),
                                      __fadd_ru(__double2float_ru(shl.hi2), eb), 0.f);
                bx0 = fmaxf(bx0, __fsub_rd(__fsub_rd(__double2float_rd(dx), rb), err));
                by0 = fmaxf(by0, __fsub_rd(__fsub_rd(__double2float_rd(dy), rb), err));
                bz0 = fmaxf(bz0, __fsub_rd(__fsub_rd(__double2float_rd(dz), rb), err));
                bx1 = fminf(bx1, __fadd_ru(__fadd_ru(__double2float_ru(dx), rb), err));
                by1 = fminf(by1, __fadd_ru(__fadd_ru(__double2float_ru(dy), rb), err));
                bz1 = fminf(bz1, __fadd_ru(__fadd_ru(__double2float_ru(dz), rb), err));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                bx0 = fmaxf(bx0, __shfl_xor_sync(FULL, bx0, o)); bx1 = fminf(bx1, __shfl_xor_sync(FULL, bx1, o));
                by0 = fmaxf(by0, __shfl_xor_sync(FULL, by0, o)); by1 = fminf(by1, __shfl_xor_sync(FULL, by1, o));
                bz0 = fmaxf(bz0, __shfl_xor_sync(FULL, bz0, o)); bz1 = fminf(bz1, __shfl_xor_sync(FULL, bz1, o));
            }
            tab_lv = lv;
            __syncwarp();
            return by0 <= by1 && bz0 <= bz1 && bx0 <= bx1;
        };

        // open level `lv`: its candidate list, narrowed to the x range of the join box (k_xindex)
        auto enter = [&](const int lv) {
            const long long lidx = (long long)h0.x * A.n_specs + rec->spec[lv];
            const CandHead hd = A.head[lidx];
            int cnt = hd.cnt;
            const long long coff = hd.off;
            if (coff + cnt > A.cand_cap) cnt = (int)max(0ll, A.cand_cap - coff);   // overflowing chunk: stay in bounds
            int s0 = 0, s1 = cnt;
            {
                if (lane == 0) n_cand += (u64)cnt;                                 // algorithm-independent work counter
                float bx0, bx1;
                const bool boxed = tables(lv, bx0, bx1);
                if (!boxed) s1 = 0;
                else if (cnt > 64 && cnt <= 65535) {
                    const unsigned short *tab = A.xtab + lidx * kXBuckets;
                    const int c0 = x_bucket(bx0, hd.x0, hd.inv_w), c1 = x_bucket(bx1, hd.x0, hd.inv_w);
                    s0 = min((int)tab[c0], cnt);
                    s1 = c1 == kXBuckets - 1 ? cnt : min(max((int)tab[c1 + 1], s0), cnt);
                }
            }
            if (lane == 0) { W.next[lv] = s0; W.end[lv] = s1; W.qn[lv] = 0; W.qi[lv] = 0; W.off[lv] = coff; }
            __syncwarp();
        };

        // emits the placed atoms 0 .. n-1 as a complete tuple
        auto emit = [&]() {
            u64 slot = 0;
            if (lane == 0) {
                slot = atomicAdd(&A.ctr->tuple_slots, 1ull);
                atomicAdd(&A.ctr->tuples, 1ull);
                atomicAdd(&A.ctr->tuple_ints, (u64)n);
            }
            slot = __shfl_sync(FULL, slot, 0);
            if ((long long)slot < A.tuple_cap) {
                int *dst = A.tuples + (long long)slot * A.tuple_stride;
                if (lane == 0) dst[0] = (int)lp;
                for (int j = lane; j < n; j += 32) dst[1 + j] = W.atom[j];
            }
            __syncwarp();
        };

        // depth-first walk of the subtree below the `base` placed atoms W.atom[0 .. base-1]
        auto walk = [&](const int base) {
        l = base;
        enter(base);
        while (true) {
            if (W.qi[l] < W.qn[l]) {
                // ---- descend into the next accepted record of this level ----
                const int a = W.queue[l][W.qi[l]];
                __syncwarp();
                if (lane == 0) { W.qi[l]++; W.atom[l] = a; W.count[l + 1 == n ? kMaxN + 1 : l + 1]++; }
                __syncwarp();
                if (l + 1 == n) {            // a complete tuple
                    emit();
                    continue;
                }
                l++;
                enter(l);
                continue;
            }
            if (W.next[l] < W.end[l]) {
                // ---- scan the next 32 records of this level's list ----
                if (tab_lv != l) {           // tables were overwritten by a deeper level: rebuild (same atoms, same values)
                    float bx0, bx1;
                    tables(l, bx0, bx1);
                }
                const long long coff = W.off[l];
                const int pos = W.next[l] + lane;
                bool ok = pos < W.end[l];
                __syncwarp();
                if (lane == 0) { W.next[l] += 32; W.qn[l] = 0; W.qi[l] = 0; }
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) r = __ldg(&A.cand[coff + pos]);
                const int a = __float_as_int(r.w);
                bool unsure = false;
                if (ok) {
                    for (int j = 0; j < l; j++) {
                        const float4 c = W.cf[j], t = W.th[j];
                        const float dx = __fsub_rn(c.x, r.x), dy = __fsub_rn(c.y, r.y), dz = __fsub_rn(c.z, r.z);
                        const float s = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                        if (s < c.w || s > t.z) { ok = false; break; }
                        unsure = unsure || !(s >= t.x && s <= t.y);
                    }
                }
                if (ok && unsure) {          // _Annulus_po on the FP64 records (eff:Annulus.h:54-68)
                    const AtomRec p = A.atoms[h1.x + a];
                    for (int j = 0; j < l; j++) {
                        const AtomRec q = A.atoms[h1.x + W.atom[j]];
                        const double t0 = __dsub_rn(q.x, p.x), t1 = __dsub_rn(q.y, p.y), t2 = __dsub_rn(q.z, p.z);
                        const double s = __dadd_rn(__dadd_rn(__dmul_rn(t0, t0), __dmul_rn(t1, t1)), __dmul_rn(t2, t2));
                        if (s < W.clo2[j] || s > W.chi2[j]) { ok = false; break; }
                    }
                }
                if (ok) {                    // TessTemplate_check (eff:TessTemplate.c:78-119)
                    const int2 rc = *reinterpret_cast<const int2 *>(&A.atoms[h1.x + a].resseq);   // resseq, chain
                    for (int j = 0; j < l; j++) {
                        const int f = W.cflag[j];
                        if (!pair_consistent(A.ignore_chain, (f & 0xff) == rc.y, (f & 0x100) != 0, W.cres[j] == rc.x,
                                             (f & 0x200) != 0)) { ok = false; break; }
                    }
                }
                const unsigned okm = __ballot_sync(FULL, ok);
                if (ok) W.queue[l][__popc(okm & lt)] = a;
                __syncwarp();
                if (lane == 0) W.qn[l] = __popc(okm);
                __syncwarp();
                continue;
            }
            // ---- this level is exhausted: back to the previous one ----
            if (l == base) break;
            l--;
        }
        __syncwarp();
        };

        if (TAIL) {
            __syncwarp();
            for (int j = lane; j < l0; j += 32) W.atom[j] = A.in[item * (l0 + 1) + 1 + j];
            __syncwarp();
            walk(l0);
            continue;
        }
        // the roots of this slice: records slice, slice + kRootSlices, ... of the first atom's list
        const CandHead hd0 = A.head[(long long)h0.x * A.n_specs + rec->spec[0]];
        int cnt0 = hd0.cnt;
        if (hd0.off + cnt0 > A.cand_cap) cnt0 = (int)max(0ll, A.cand_cap - hd0.off);
        for (int root = slice; root < cnt0; root += kRootSlices) {
            const int a0 = __float_as_int(__ldg(&A.cand[hd0.off + root].w));
            __syncwarp();
            if (lane == 0) { W.atom[0] = a0; if (n > 1) W.count[1]++; }
            __syncwarp();
            if (n == 1) { emit(); continue; }
            walk(1);
        }
    }
    __syncwarp();
    // work counters: F_l of every size, tuples' ints, candidate tests
    for (int i = 1 + lane; i <= kMaxN; i += 32)
        if (W.count[i]) atomicAdd(&A.ctr->level[i], (u64)W.count[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_cand += __shfl_down_sync(FULL, n_cand, o);
    if (lane == 0 && n_cand) atomicAdd(&A.ctr->cand_tests, n_cand);
}

}  // namespace jb
