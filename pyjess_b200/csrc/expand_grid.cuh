// expand_grid.cuh -- frontier expansion over the cell grids of the candidate lists.
//
// k_expand_lv (expand_lv.cuh) answers "which records of the list lie inside the join box" by
// streaming the x range of the box through six FP32 compares per record: 23 box tests per
// partial match on BASELINE config 3, 51 on config 4, and more than half of its issue slots go
// to the bookkeeping around that stream (chunk numbering, owner look-up, survivor queue;
// profiles/r1ae).  k_expand_grid<L, G, ROOT> keeps everything around the range query -- the
// staging of the placed atoms, the outward-rounded FP32 join box (Join_computeBox,
// eff:Join.c:82-111), the exact FP64 drain (_Annulus_po eff:Annulus.h:54-68 +
// TessTemplate_check eff:TessTemplate.c:78-119), the emission -- and replaces the stream by the
// cell grid of kernels.cuh (BinTab): per bin of 64 records, three slab look-ups
//
//     survivors = AND_a  upto[a][cell(hi_a)] & ~upto[a][cell(lo_a) - 1]
//
// i.e. six 8-byte loads and a handful of logic operations per (partial match, bin), all in the
// lanes that own the partial match: no chunk table, no owner search, no per-record compare.
// The survivors are then handed to the exact test one per lane: a warp prefix sum over the popcounts
// numbers them, lane t of a round looks up the owner of survivor 32r + t (binary search over the
// prefix sums, by shuffles) and the position of its bit (select over the owner's mask words) --
// a fixed cost per 32 survivors however they are spread over the partial matches.  Lists longer than one bin are x-sorted, so
// the bins that can meet the box are the ones of its x range (the list's x-bucket index).
//
// The masks are a superset of the box content (monotone cell function), the box is a superset of
// the shell intersection: reference semantics are untouched, the FP64 predicate decides.
#pragma once
#include "kernels.cuh"

namespace jb {

// position of the k-th (0-based) set bit of x; k < popc(x)
__device__ __forceinline__ int nth_set_bit(unsigned x, int k)
{
    int pos = 0, t;
    t = __popc(x & 0xffffu); if (k >= t) { k -= t; pos += 16; x >>= 16; }
    t = __popc(x & 0xffu);   if (k >= t) { k -= t; pos += 8;  x >>= 8; }
    t = __popc(x & 0xfu);    if (k >= t) { k -= t; pos += 4;  x >>= 4; }
    t = __popc(x & 0x3u);    if (k >= t) { k -= t; pos += 2;  x >>= 2; }
    if (k >= (int)(x & 1u)) pos += 1;
    return pos;
}

constexpr int kGridQueue = 512;                         // survivors per window of the per-warp queue
static_assert(kBinRecs <= 128, "queue entries hold 7 position bits");

template <int L, int G> struct alignas(16) GridWarp {   // per-warp shared memory
    static constexpr int IPW = 32 / G;                  // partial matches per batch
    // placed atoms + shells towards the new atom, one array per field and [atom][partial match].  The FP32
    // image decides all but the borderline cases (see the drain), the FP64 radii serve those.
    float4 cf[L][IPW];                                  // centre relative to the structure origin (FP32), lo_reject
    float4 th[L][IPW];                                  // lo_certain, hi_certain, hi_reject of the squared distance
    // (the FP64 radii are NOT staged: the few records per million inside the error band read them from the shell
    //  table again -- 16 bytes per placed atom less, which is what lets level 3 keep four CTAs per SM)
    int cres[L][IPW], cflag[L][IPW];
    unsigned cbase[IPW];                                // first record of the candidate list (capi.cu bounds both tables by 2^32)
    unsigned abase[IPW];                                // first atom of the structure
    int2 meta[IPW];                                     // pair index, "this level completes the template"
                                      // records in the candidate list
    int chain_atom[L][IPW];
    // deep levels (24-atom templates): half the staging, so that two CTAs still share an SM
    static constexpr int kStage = L > 12 ? kStageItems / 2 : kStageItems;
    int stage[kStage * (L + 2)];                        // accepted extensions awaiting the flush
    unsigned short queue[kGridQueue];                   // box survivors of the current window: (owner lane << 7) | position in its bin
    long long t_next;                                   // the warp's chunk of the tuple list: next free record, records left,
    int t_left;                                         // tuples written so far (flushed into the counters at the end),
    unsigned n_tup;                                     // size of the next chunk (8 doubling to kTupleChunk: a warp that
    int t_chunk, t_pad[3];                              // completes a handful of tuples leaves a handful of holes)
};

template <int L, int G> constexpr size_t expand_grid_smem() { return sizeof(GridWarp<L, G>) * kExpandWarps; }

struct GridArgs {
    const BinTab *bins;
    long long bin_cap;
};

template <int L, int G, bool ROOT = false>
__global__ void __launch_bounds__(kExpandThreads, 4) k_expand_grid(const ExpandArgs A, const GridArgs B, const int stats)
{
    static_assert(!ROOT || L == 1, "root mode feeds level 1");
    constexpr int IPW = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int in_stride = L + 1, out_stride = L + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GridWarp<L, G> &W = reinterpret_cast<GridWarp<L, G> *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const unsigned lt = (1u << lane) - 1;
    int fill = 0;
    if (lane == 0) { W.t_next = 0; W.t_left = 0; W.n_tup = 0u; W.t_chunk = 8; }
    __syncwarp();

    if (!ROOT && A.ctr->tail_level != 0 && A.ctr->tail_level <= (u64)L) return;   // finished depth first (k_expand_tail)
    u64 n_in = *A.in_count;
    if ((long long)n_in > A.in_cap) n_in = (u64)A.in_cap;
    u64 n_cand = 0;                                   // per-lane partial sums
    unsigned n_scan = 0, n_box = 0;                   // per-lane partial sums (diagnostics)
    const float INF = __int_as_float(0x7f800000);

    // Exact predicate + emission: one box survivor per lane (record `pos` of the list of partial match g).
    // _Annulus_po (eff:Annulus.h:54-68) decides on s = ((dx^2 + dy^2) + dz^2) in FP64.  The FP32 image of
    // the same difference (coordinates relative to the structure origin, |error of s| < e with
    // e = 2^-20 rb (ext + rb), more than twice the worst case: DESIGN.md "FP32 first") settles every
    // survivor whose s lies outside [lo2 - e, lo2 + e] and [hi2 - e, hi2 + e]; the others -- a few per
    // million -- take the FP64 operations of the reference on the FP64 atom records.
    auto drain = [&](bool ok, const float4 r, const int g) {
        const int a = __float_as_int(r.w);
        const int2 mt = W.meta[g];
        bool unsure = false;
        if (ok) {
#pragma unroll
            for (int j = 0; j < L; j++) {
                const float4 c = W.cf[j][g], t = W.th[j][g];
                const float dx = __fsub_rn(c.x, r.x), dy = __fsub_rn(c.y, r.y), dz = __fsub_rn(c.z, r.z);
                const float s = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                if (s < c.w || s > t.z) { ok = false; break; }
                unsure = unsure || !(s >= t.x && s <= t.y);
            }
        }
        if (__any_sync(FULL, ok && unsure)) {
            if (ok && unsure) {
                const AtomRec p = A.atoms[W.abase[g] + a];
                const PairRec *rec = A.recs + mt.x;
                const int n = rec->n, k = rec->order[L];
#pragma unroll
                for (int j = 0; j < L; j++) {
                    const AtomRec q = A.atoms[W.abase[g] + W.chain_atom[j][g]];
                    const ShellRec *sh = A.shell + rec->shell0 + (long long)rec->order[j] * n + k;
                    const double t0 = __dsub_rn(q.x, p.x), t1 = __dsub_rn(q.y, p.y), t2 = __dsub_rn(q.z, p.z);
                    const double s = __dadd_rn(__dadd_rn(__dmul_rn(t0, t0), __dmul_rn(t1, t1)), __dmul_rn(t2, t2));
                    if (s < sh->lo2 || s > sh->hi2) { ok = false; break; }
                }
            }
        }
        if (ok) {
            const int2 rc = *reinterpret_cast<const int2 *>(&A.atoms[W.abase[g] + a].resseq);   // resseq, chain
#pragma unroll
            for (int j = 0; j < L; j++) {
                const int f = W.cflag[j][g];
                if (!pair_consistent(A.ignore_chain, (f & 0xff) == rc.y, (f & 0x100) != 0, W.cres[j][g] == rc.x,
                                     (f & 0x200) != 0)) { ok = false; break; }
            }
        }
        const unsigned okm = __ballot_sync(FULL, ok);
        if (okm == 0) return;
        const unsigned cm = __ballot_sync(FULL, ok && mt.y);
        const unsigned nm = okm & ~cm;
        if (cm) {
            // Finished tuples go straight to the tuple list.  Its records are reserved kTupleChunk at a time per warp: one
            // atomic per round of 32 survivors (two, with the statistics) on ONE address serialised the completing levels
            // of dense-hit workloads -- 8 M tuples at level 8 of cfg3dense cost 2.2 of its 4.8 ms (gpurun_out/r3n).  A
            // round that does not fit the rest of the chunk fills it and continues in a new one; only the last chunk of a
            // warp is left partly unused, its records marked as holes at the end of the kernel.
            const int k = __popc(cm);
            const long long t_next = W.t_next;
            const int t_left = W.t_left;
            long long fresh = 0;
            const int want = max(W.t_chunk, k - t_left);     // (only used when the round does not fit)
            if (k > t_left) {
                u64 v = 0;
                if (lane == 0) v = atomicAdd(&A.ctr->tuple_slots, (u64)want);
                fresh = (long long)__shfl_sync(FULL, v, 0);
            }
            if (ok && mt.y) {
                const int r = __popc(cm & lt);
                const long long slot = r < t_left ? t_next + r : fresh + (r - t_left);
                if (slot < A.tuple_cap) {
                    int *dst = A.tuples + slot * A.tuple_stride;
                    dst[0] = mt.x;
#pragma unroll
                    for (int j = 0; j < L; j++) dst[1 + j] = W.chain_atom[j][g];
                    dst[1 + L] = a;
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (k > t_left) {
                    W.t_next = fresh + (k - t_left); W.t_left = want - (k - t_left);
                    W.t_chunk = min(kTupleChunk, 2 * W.t_chunk);
                }
                else { W.t_next = t_next + k; W.t_left = t_left - k; }
                W.n_tup += (unsigned)k;
            }
            __syncwarp();
        }
        if (nm) {
            if (ok && !mt.y) {
                int *dst = W.stage + (fill + __popc(nm & lt)) * out_stride;
                dst[0] = mt.x;
#pragma unroll
                for (int j = 0; j < L; j++) dst[1 + j] = W.chain_atom[j][g];
                dst[1 + L] = a;
            }
            fill += __popc(nm);
            if (fill + 32 > GridWarp<L, G>::kStage) {   // flush
                __syncwarp();
                u64 ob = 0;
                if (lane == 0) ob = atomicAdd(A.out_count, (u64)fill);
                ob = __shfl_sync(FULL, ob, 0);
                if ((long long)ob < A.out_cap) {   // on overflow write the part that fits: every slot below
                    const int fit = (int)min((long long)fill, A.out_cap - (long long)ob);   // the capacity holds a real item
                    int *dst = A.out + (long long)ob * out_stride;
                    for (int i = lane; i < fit * out_stride; i += 32) dst[i] = W.stage[i];
                }
                fill = 0;
                __syncwarp();
            }
        }
    };

    // Hand the survivors of the batch's current bins -- the set bits of the masks held by the first
    // lane of every group -- to the exact test, 32 per round.  A loose box passes a whole bin, a tight
    // one nothing: one lane testing its own bits ran 65-90 rounds per batch at 4 active lanes
    // (gpurun_out/r2d-r2f), and a per-round owner search by shuffles cost ~55 instructions per 32
    // survivors (30 % of the kernel, profiles/r2i).  Instead every lane only UNPACKS its bits -- a
    // few instructions per survivor -- into a per-warp queue of (owner lane, position) entries at
    // the slots a warp prefix sum assigns, and the rounds read the queue: survivor s of a window is
    // queue[s], whoever owns it.  A window holds kGridQueue survivors; lanes whose slots lie beyond
    // it keep their remaining bits for the next window.
    auto deal = [&](BinMask &m, const int posbase) {
        int c = 0;
#pragma unroll
        for (int h = 0; h < kMaskWords; h++) c += __popc(m.w[h]);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const int total = __shfl_sync(FULL, inc, 31);
        int idx = inc - c;                                   // slot of this lane's next survivor
        for (int w0 = 0; w0 < total; w0 += kGridQueue) {
            const int w1 = min(total, w0 + kGridQueue);
#pragma unroll
            for (int h = 0; h < kMaskWords; h++) {
                unsigned w = m.w[h];
                while (w != 0u && idx < w1) {
                    const int b = __ffs(w) - 1;
                    w &= w - 1u;
                    W.queue[idx - w0] = (unsigned short)((lane << 7) | (32 * h + b));
                    idx++;
                }
                m.w[h] = w;
            }
            __syncwarp();
            for (int s0 = w0; s0 < w1; s0 += 32) {
                const int s = s0 + lane;
                const bool ok = s < w1;
                const int e = ok ? (int)W.queue[s - w0] : 0;
                const int o = e >> 7;
                const int pb = __shfl_sync(FULL, posbase, o);
                const int g = o / G;
                // (prefetching the record of round r + 1 while round r is tested measured neutral: 8.01 -> 8.00 ms, r3c)
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) r = __ldg(&A.cand[W.cbase[g] + pb + (e & 127)]);
                drain(ok, r, g);
            }
            __syncwarp();                                    // the next window overwrites the queue
        }
    };

    // Work is CLAIMED, not dealt: a warp takes the next kGrab batches (pairs in root mode) off a device counter when it
    // runs dry.  The static round-robin deal of round 1 left the SMs idle for 10-24 % of a launch -- partial matches of
    // large structures cost several times the average, and the warps that drew them finished last (profiles/r2_sweep_ncu.md:
    // sm__cycles_active 76 / 87 / 90 % of the elapsed cycles at levels 1 / 2 / 3).
    // (a claim is 2 batches / 4 pairs; one when the level holds too little work to keep every warp busy that way)
    const int grab = (ROOT ? 1 : IPW) * ((long long)n_in >= (long long)gridDim.x * kExpandWarps * (ROOT ? 8 : 4 * IPW) ? (ROOT ? 4 : 2) : 1);
    u64 *const sched = &A.ctr->sched[L];
    long long claim = 0, claim_end = 0;                                   // warp uniform
    auto claim_more = [&]() {
        u64 g0 = 0;
        if (lane == 0) g0 = atomicAdd(sched, (u64)grab);
        claim = (long long)__shfl_sync(FULL, g0, 0);
        claim_end = claim + grab;
    };
    claim_more();
    long long my_pair = claim;                                            // ROOT: next pair of this warp
    int root_j = 0, root_cnt = 0;
    long long root_off = 0;
    bool pair_loaded = false;
    unsigned n_roots = 0;
    for (;;) {
        long long base_item = 0;
        if (!ROOT) {
            if (claim >= claim_end) claim_more();
            base_item = claim;
            claim += IPW;
            if (base_item >= (long long)n_in) break;
        }
        const long long item = base_item + grp;
        bool valid = item < (long long)n_in;
        int root_lp = 0, root_a = 0;
        if (ROOT) {
            int filled = 0;
            valid = false;
            while (filled < IPW && my_pair < (long long)n_in) {            // all conditions are warp uniform
                if (!pair_loaded) {
                    const PairRec *rec = A.recs + my_pair;
                    const CandHead hd = A.head[(long long)rec->m * A.n_specs + rec->spec[0]];
                    root_cnt = hd.cnt; root_off = hd.off; root_j = 0;
                    if (root_off + root_cnt > A.cand_cap) root_cnt = (int)max(0ll, A.cand_cap - root_off);
                    pair_loaded = true;
                }
                const int take = min(root_cnt - root_j, IPW - filled);
                if (grp >= filled && grp < filled + take) {
                    valid = true;
                    root_lp = (int)my_pair;
                    root_a = __float_as_int(__ldg(&A.cand[root_off + root_j + (grp - filled)].w));
                }
                filled += take;
                root_j += take;
                if (root_j >= root_cnt) {
                    pair_loaded = false;
                    if (++my_pair >= claim_end) { claim_more(); my_pair = claim; }
                }
            }
            if (filled == 0) break;
            if (valid && gl == 0) n_roots++;
        }
        int cnt = 0, bin0 = 0;
        long long coff = 0, lidx = 0;
        float hx0 = 0.f, hinv = 0.f;
        float bx0 = -INF, by0 = -INF, bz0 = -INF, bx1 = INF, by1 = INF, bz1 = INF;
        if (valid) {
            const int *it = ROOT ? nullptr : A.in + item * in_stride;
            const int lp = ROOT ? root_lp : it[0];
            const PairRec *rec = A.recs + lp;
            const int4 h0 = *reinterpret_cast<const int4 *>(rec);            // m, t, ta0, n
            const longlong2 h1 = *reinterpret_cast<const longlong2 *>(&rec->abase);
            const int n = h0.w;
            const int k = rec->order[L];
            lidx = (long long)h0.x * A.n_specs + rec->spec[L];
            const CandHead hd = A.head[lidx];
            cnt = hd.cnt;
            coff = hd.off;
            bin0 = hd.bin0;
            if (coff + cnt > A.cand_cap) cnt = (int)max(0ll, A.cand_cap - coff);   // overflowing chunk: stay in bounds
            hx0 = hd.x0; hinv = hd.inv_w;
            if (gl == 0) {
                W.meta[grp] = make_int2(lp, (L + 1 == n) ? 1 : 0);
                W.abase[grp] = (unsigned)h1.x;
                W.cbase[grp] = (unsigned)coff;
            }
            const MolExt me = A.ext[h0.x];
            const float err = __double2float_ru(me.ext * 2.4e-7 + 1e-30);    // > ext * 2^-22: record rounding
            const float extu = __double2float_ru(me.ext);
#pragma unroll
            for (int jj = 0; jj < (L + G - 1) / G; jj++) {
                const int j = jj * G + gl;
                if (j < L) {
                    const int a = ROOT ? root_a : it[1 + j];
                    const AtomRec ar = A.atoms[h1.x + a];
                    const ShellRec sh = A.shell[h1.y + (long long)rec->order[j] * n + k];
                    W.cres[j][grp] = ar.resseq; W.cflag[j][grp] = ar.chain | (sh.flags << 8);
                    W.chain_atom[j][grp] = a;
                    const double dx = __dsub_rn(ar.x, me.ox), dy = __dsub_rn(ar.y, me.oy), dz = __dsub_rn(ar.z, me.oz);
                    const float rb = sh.rb;
                    // FP32 image for the drain: the centre exactly as k_sort_atoms rounds the records, and the
                    // squared radii moved outwards / inwards by the error bound of the FP32 squared distance
                    const float eb = __fmul_ru(__fmul_ru(9.5367431640625e-07f, rb), __fadd_ru(extu, rb));   // 2^-20 rb (ext + rb)
                    W.cf[j][grp] = make_float4(__double2float_rn(dx), __double2float_rn(dy), __double2float_rn(dz),
                                               __fsub_rd(__double2float_rd(sh.lo2), eb));
                    W.th[j][grp] = make_float4(__fadd_ru(__double2float_ru(sh.lo2), eb), __fsub_rd(__double2float_rd(sh.hi2), eb),
                                               __fadd_ru(__double2float_ru(sh.hi2), eb), 0.f);
                    bx0 = fmaxf(bx0, __fsub_rd(__fsub_rd(__double2float_rd(dx), rb), err));
                    by0 = fmaxf(by0, __fsub_rd(__fsub_rd(__double2float_rd(dy), rb), err));
                    bz0 = fmaxf(bz0, __fsub_rd(__fsub_rd(__double2float_rd(dz), rb), err));
                    bx1 = fminf(bx1, __fadd_ru(__fadd_ru(__double2float_ru(dx), rb), err));
                    by1 = fminf(by1, __fadd_ru(__fadd_ru(__double2float_ru(dy), rb), err));
                    bz1 = fminf(bz1, __fadd_ru(__fadd_ru(__double2float_ru(dz), rb), err));
                }
            }
            if (gl == 0) n_cand += (u64)cnt;     // algorithm-independent work counter
        }
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) {   // join box = intersection over the placed atoms
            bx0 = fmaxf(bx0, __shfl_xor_sync(FULL, bx0, o)); bx1 = fminf(bx1, __shfl_xor_sync(FULL, bx1, o));
            by0 = fmaxf(by0, __shfl_xor_sync(FULL, by0, o)); by1 = fminf(by1, __shfl_xor_sync(FULL, by1, o));
            bz0 = fmaxf(bz0, __shfl_xor_sync(FULL, bz0, o)); bz1 = fminf(bz1, __shfl_xor_sync(FULL, bz1, o));
        }
        // an empty join box kills the partial match without touching its candidates
        const bool boxed = valid && cnt > 0 && bx0 <= bx1 && by0 <= by1 && bz0 <= bz1;
        // ---- bins that can meet the box: all of a short list, the x range of a long one ------------
        int b_lo = 0, b_hi = boxed ? (cnt + kBinRecs - 1) / kBinRecs : 0;
        if (boxed && cnt > kBinRecs && cnt <= 65535) {
            const unsigned short *tab = A.xtab + lidx * kXBuckets;
            const int c0 = x_bucket(bx0, hx0, hinv), c1 = x_bucket(bx1, hx0, hinv);
            const int s0 = min((int)tab[c0], cnt);                           // first record of the first bucket
            const int s1 = c1 == kXBuckets - 1 ? cnt : min(max((int)tab[c1 + 1], s0), cnt);
            b_lo = s0 / kBinRecs;
            b_hi = s1 > s0 ? (s1 + kBinRecs - 1) / kBinRecs : b_lo;
        }
        int nbv = b_hi - b_lo;
        int maxb = nbv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxb = max(maxb, __shfl_xor_sync(FULL, maxb, o));
        __syncwarp();                                    // staging of the batch is visible to the drains
        for (int bi = 0; bi < maxb; bi++) {
            BinMask m;
#pragma unroll
            for (int h = 0; h < kMaskWords; h++) m.w[h] = 0u;
            const long long bin = (long long)bin0 + b_lo + bi;
            const bool on = bi < nbv && bin < B.bin_cap;
            if (on) {
                const BinTab *T = B.bins + bin;
#pragma unroll
                for (int h = 0; h < kMaskWords; h++) m.w[h] = ~0u;
#pragma unroll
                for (int aa = 0; aa < (3 + G - 1) / G; aa++) {
                    const int ax = aa * G + gl;
                    if (ax < 3) {
                        const float lo = ax == 0 ? bx0 : (ax == 1 ? by0 : bz0), hi = ax == 0 ? bx1 : (ax == 1 ? by1 : bz1);
                        const float2 fr = __ldg(&T->frame[ax]);
                        const int cl = axis_cell(lo, fr), ch = axis_cell(hi, fr);
                        const BinMask up = ldg_mask(&T->upto[ax][ch]);
                        BinMask dn;
#pragma unroll
                        for (int h = 0; h < kMaskWords; h++) dn.w[h] = 0u;
                        if (cl > 0) dn = ldg_mask(&T->upto[ax][cl - 1]);
#pragma unroll
                        for (int h = 0; h < kMaskWords; h++) m.w[h] &= up.w[h] & ~dn.w[h];
                    }
                }
            }
            if (G > 1) {
#pragma unroll
                for (int o = G >> 1; o > 0; o >>= 1)
#pragma unroll
                    for (int h = 0; h < kMaskWords; h++) m.w[h] &= __shfl_xor_sync(FULL, m.w[h], o);
                if (gl != 0)
#pragma unroll
                    for (int h = 0; h < kMaskWords; h++) m.w[h] = 0u;
            }
            if (stats && gl == 0 && on) {
                n_scan += (unsigned)min(cnt - (b_lo + bi) * kBinRecs, kBinRecs);
#pragma unroll
                for (int h = 0; h < kMaskWords; h++) n_box += __popc(m.w[h]);
            }
            deal(m, (b_lo + bi) * kBinRecs);
        }
        __syncwarp();   // the next batch overwrites the per-warp tables
        if (stats) {
            unsigned ts = n_scan, tb = n_box;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { ts += __shfl_down_sync(FULL, ts, o); tb += __shfl_down_sync(FULL, tb, o); }
            if (lane == 0 && (ts | tb)) {
                atomicAdd(&A.ctr->scan_tests, (u64)ts);
                atomicAdd(&A.ctr->box_pass, (u64)tb);
            }
            n_scan = n_box = 0;
        }
    }
    __syncwarp();
    if (fill > 0) {
        u64 ob = 0;
        if (lane == 0) ob = atomicAdd(A.out_count, (u64)fill);
        ob = __shfl_sync(FULL, ob, 0);
        if ((long long)ob < A.out_cap) {
            const int fit = (int)min((long long)fill, A.out_cap - (long long)ob);
            int *dst = A.out + (long long)ob * out_stride;
            for (int i = lane; i < fit * out_stride; i += 32) dst[i] = W.stage[i];
        }
    }
    {   // the unused rest of the warp's last chunk of the tuple list: holes (k_superpose skips a negative pair)
        const long long t_next = W.t_next;
        const int t_left = W.t_left;
        for (int i = lane; i < t_left; i += 32)
            if (t_next + i < A.tuple_cap) A.tuples[(t_next + i) * A.tuple_stride] = -1;
        if (lane == 0 && W.n_tup) {
            atomicAdd(&A.ctr->tuples, (u64)W.n_tup);
            atomicAdd(&A.ctr->tuple_ints, (u64)W.n_tup * (u64)(L + 1));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_cand += __shfl_down_sync(FULL, n_cand, o);
    if (lane == 0 && n_cand) atomicAdd(&A.ctr->cand_tests, n_cand);
    if (ROOT) {   // the roots are the partial matches with one atom: F_1 of the work counters
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_roots += __shfl_down_sync(FULL, n_roots, o);
        if (lane == 0 && n_roots) atomicAdd(&A.ctr->level[1], (u64)n_roots);
    }
}

}  // namespace jb
