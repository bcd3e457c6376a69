// expand_lv.cuh -- level-specialised frontier expansion for the hot levels.
//
// k_expand (kernels.cuh) takes the level as a run-time argument: every shared-memory
// offset, item stride and inner-loop bound is re-derived in the hot loops (16 % of
// its warp instructions on BASELINE config 3, profiles/r1n), and its scan loop runs
// until the LONGEST of the 32/G candidate ranges of a batch is exhausted (14.1
// iterations per batch where 6.2 would do).  k_expand_lv<L, G> is the same
// algorithm -- same staging, same FP32 join box, same exact FP64 predicate, same
// emission -- specialised on the level L (atoms already placed) with
//
//   * all strides/offsets compile-time, loops over the placed atoms unrolled;
//   * BOTH x bounds of the join box located through the list's x-bucket index
//     (k_xindex: two independent 2-byte loads instead of two chains of 5-6 dependent
//     probes; a short binary search only refines buckets holding > 4 records), so
//     that every partial match owns a record range (the k-d tree walk of
//     eff:KdTree.c:238-296 becomes a range);
//   * a LOAD-BALANCED, COALESCED scan: the ranges of a batch are cut into chunks of
//     kChunk records, numbered by a warp prefix sum and dealt to QUADS of lanes
//     whatever partial match they belong to (owner found by a branch-free binary
//     search over the chunk offsets in shared memory).  A quad reads 64 contiguous
//     bytes per load instruction: profiles/r1t showed the lane-per-chunk layout
//     (32 different lines per LDG) saturating the L1 data stage at 90-93 %.
//
// Reference semantics are untouched: the decision is still _Annulus_po in FP64
// with explicit RN operations (eff:Annulus.h:54-68) plus TessTemplate_check
// (eff:TessTemplate.c:78-119); FP32 only ever rejects records that are provably
// outside the outward-rounded join box.
#pragma once
#include "kernels.cuh"

namespace jb {

constexpr int kMaxLv = 12;       // levels 1..kMaxLv have a specialisation; deeper levels use k_expand
#ifndef JB_CHUNK
#define JB_CHUNK 16
#endif
#ifndef JB_CHUNK_L1
#define JB_CHUNK_L1 JB_CHUNK
#endif
#ifndef JB_CHUNK_DEEP
#define JB_CHUNK_DEEP JB_CHUNK
#endif
#ifndef JB_DEEP_FROM
#define JB_DEEP_FROM 99
#endif
// records per scan chunk, read by a QUAD of lanes (64-byte pieces); per level: the single-shell box of
// level 1 leaves long x ranges, the join boxes of deep levels short ones
template <int L> struct ChunkOf { static constexpr int value = L == 1 ? JB_CHUNK_L1 : (L >= JB_DEEP_FROM ? JB_CHUNK_DEEP : JB_CHUNK); };
constexpr int kOwnerCap = 128;   // batches of up to this many chunks get a direct chunk -> owner table
constexpr int kQuad = 4;         // lanes per chunk: one LDG.128 of a quad covers 64 contiguous bytes

template <int L, int G> struct alignas(16) LvWarp {   // per-warp shared memory
    static constexpr int IPW = 32 / G;                // partial matches per batch
    // placed atoms + shells towards the new atom, one array per field and [atom][partial match]: the drain's
    // lanes read the records of DIFFERENT partial matches, which the 48-byte ChainRec layout served with
    // 2- to 16-way bank conflicts depending on L (stride L x 12 words)
    double cx[L][IPW], cy[L][IPW], cz[L][IPW], clo2[L][IPW], chi2[L][IPW];
    int cres[L][IPW], cflag[L][IPW];
    float4 blo[IPW], bhi[IPW];                        // join box (outward rounded)
    long long cbase[IPW];                             // first record of the x range
    long long abase[IPW];                             // first atom of the structure
    int2 meta[IPW];                                   // pair index, "this level completes the template"
    int clen[IPW];                                    // records in the x range
    int cstart[IPW];                                  // first chunk of the partial match (exclusive scan)
    int2 queue[32 + 8 * ChunkOf<L>::value];               // box survivors (atom, partial match): < 32 carried + 8 chunks new
    unsigned char owner[kOwnerCap];                   // chunk -> partial match of the batch (short batches)
    int chain_atom[L][IPW];
    int stage[kStageItems * (L + 2)];                 // accepted extensions awaiting the flush
};

template <int L, int G> constexpr size_t expand_lv_smem() { return sizeof(LvWarp<L, G>) * kExpandWarps; }

// ROOT (level 1 only): the level-0 launch and the level-1 frontier disappear -- every warp walks
// its own live pairs (pair = warp id, + number of warps, ...) and deals the candidates of the first
// scan atom straight into its batches, whatever pair they belong to.  A.in_count = live pairs.
template <int L, int G, bool ROOT = false>
__global__ void __launch_bounds__(kExpandThreads, 4) k_expand_lv(const ExpandArgs A, const int stats)
{
    static_assert(!ROOT || L == 1, "root mode feeds level 1");
    constexpr int IPW = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int in_stride = L + 1, out_stride = L + 2;
    constexpr int kChunk = ChunkOf<L>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LvWarp<L, G> &W = reinterpret_cast<LvWarp<L, G> *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const unsigned lt = (1u << lane) - 1;
    int fill = 0, qfill = 0;

    u64 n_in = *A.in_count;
    if ((long long)n_in > A.in_cap) n_in = (u64)A.in_cap;
    const long long total_warps = (long long)gridDim.x * kExpandWarps;
    u64 n_cand = 0;                                   // per-lane partial sums
    unsigned n_scan = 0, n_box = 0;                   // lane 0 only (diagnostics)
    const float INF = __int_as_float(0x7f800000);

    // exact predicate + emission for queue entries [q0, q0 + nq)
    auto drain = [&](int q0, int nq) {
        __syncwarp();
        bool ok = lane < nq;
        int a = 0, g = 0;
        if (ok) {
            const int2 e = W.queue[q0 + lane];
            a = e.x; g = e.y;
        }
        __syncwarp();   // the scan loop reuses these queue slots right after the drain (racecheck: WAR)
        const int2 mt = W.meta[g];
        if (ok) {
            const AtomRec p = A.atoms[W.abase[g] + a];
            // _Annulus_po (eff:Annulus.h:54-68): ((dx^2 + dy^2) + dz^2), centre - point
#pragma unroll
            for (int j = 0; j < L; j++) {
                const double t0 = __dsub_rn(W.cx[j][g], p.x), t1 = __dsub_rn(W.cy[j][g], p.y), t2 = __dsub_rn(W.cz[j][g], p.z);
                const double s = __dadd_rn(__dadd_rn(__dmul_rn(t0, t0), __dmul_rn(t1, t1)), __dmul_rn(t2, t2));
                if (s < W.clo2[j][g] || s > W.chi2[j][g]) { ok = false; break; }
            }
            if (ok) {
#pragma unroll
                for (int j = 0; j < L; j++) {
                    const int f = W.cflag[j][g];
                    if (!pair_consistent(A.ignore_chain, (f & 0xff) == p.chain, (f & 0x100) != 0, W.cres[j][g] == p.resseq,
                                         (f & 0x200) != 0)) { ok = false; break; }
                }
            }
        }
        const unsigned okm = __ballot_sync(FULL, ok);
        if (okm == 0) return;
        const unsigned cm = __ballot_sync(FULL, ok && mt.y);
        const unsigned nm = okm & ~cm;
        if (cm) {   // finished tuples go straight to the tuple list (one atomic per warp)
            u64 tb = 0;
            if (lane == 0) {
                tb = atomicAdd(&A.ctr->tuple_slots, (u64)__popc(cm));
                atomicAdd(&A.ctr->tuples, (u64)__popc(cm));
                atomicAdd(&A.ctr->tuple_ints, (u64)__popc(cm) * (u64)(L + 1));
            }
            tb = __shfl_sync(FULL, tb, 0);
            if (ok && mt.y) {
                const long long slot = (long long)tb + __popc(cm & lt);
                if (slot < A.tuple_cap) {
                    int *dst = A.tuples + slot * A.tuple_stride;
                    dst[0] = mt.x;
#pragma unroll
                    for (int j = 0; j < L; j++) dst[1 + j] = W.chain_atom[j][g];
                    dst[1 + L] = a;
                }
            }
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
            if (fill + 32 > kStageItems) {   // flush
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

    long long my_pair = (long long)blockIdx.x * kExpandWarps + warp;      // ROOT: next pair of this warp
    int root_j = 0, root_cnt = 0;
    long long root_off = 0;
    bool pair_loaded = false;
    unsigned n_roots = 0;
    for (long long base_item = ((long long)blockIdx.x * kExpandWarps + warp) * IPW; ROOT || base_item < (long long)n_in;
         base_item += total_warps * IPW) {
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
                if (root_j >= root_cnt) { my_pair += total_warps; pair_loaded = false; }
            }
            if (filled == 0) break;
            if (valid && gl == 0) n_roots++;
        }
        int cnt = 0;
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
            if (coff + cnt > A.cand_cap) cnt = (int)max(0ll, A.cand_cap - coff);   // overflowing chunk: stay in bounds
            hx0 = hd.x0; hinv = hd.inv_w;
            if (gl == 0) {
                W.meta[grp] = make_int2(lp, (L + 1 == n) ? 1 : 0);
                W.abase[grp] = h1.x;
            }
            const MolExt me = A.ext[h0.x];
            const float err = __double2float_ru(me.ext * 2.4e-7 + 1e-30);    // > ext * 2^-22: record rounding
#pragma unroll
            for (int jj = 0; jj < (L + G - 1) / G; jj++) {
                const int j = jj * G + gl;
                if (j < L) {
                    const int a = ROOT ? root_a : it[1 + j];
                    const AtomRec ar = A.atoms[h1.x + a];
                    const ShellRec sh = A.shell[h1.y + (long long)rec->order[j] * n + k];
                    W.cx[j][grp] = ar.x; W.cy[j][grp] = ar.y; W.cz[j][grp] = ar.z;
                    W.clo2[j][grp] = sh.lo2; W.chi2[j][grp] = sh.hi2;
                    W.cres[j][grp] = ar.resseq; W.cflag[j][grp] = ar.chain | (sh.flags << 8);
                    W.chain_atom[j][grp] = a;
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
            if (gl == 0) n_cand += (u64)cnt;     // algorithm-independent work counter
        }
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) {   // join box = intersection over the placed atoms
            bx0 = fmaxf(bx0, __shfl_xor_sync(FULL, bx0, o)); bx1 = fminf(bx1, __shfl_xor_sync(FULL, bx1, o));
            by0 = fmaxf(by0, __shfl_xor_sync(FULL, by0, o)); by1 = fminf(by1, __shfl_xor_sync(FULL, by1, o));
            bz0 = fmaxf(bz0, __shfl_xor_sync(FULL, bz0, o)); bz1 = fminf(bz1, __shfl_xor_sync(FULL, bz1, o));
        }
        // an empty join box kills the partial match without touching its candidates
        const bool boxed = valid && bx0 <= bx1 && by0 <= by1 && bz0 <= bz1;
        // ---- x range of the box in the x-sorted list: two binary searches, stopped at a span
        //      of two records (the scan re-tests x, so a slightly wide range is harmless)
        int lo = 0, hi = cnt;
        const bool narrow = boxed && cnt > 2 * kChunk && cnt <= 65535;
        {
            // searcher s of the group: s == 0 first record with x >= bx0, s == 1 first record with x > bx1
            // (G == 1: the lane runs both searches one after the other)
            const unsigned short *tab = A.xtab + ((long long)lidx) * kXBuckets;
#pragma unroll
            for (int pass = 0; pass < (G == 1 ? 2 : 1); pass++) {
                const int which = G == 1 ? pass : gl;
                const bool searcher = narrow && which < 2;
                const float key = which == 0 ? bx0 : bx1;
                int s0 = 0, s1 = cnt;
                if (searcher) {
                    const int b = x_bucket(key, hx0, hinv);
                    s0 = min((int)tab[b], cnt);
                    s1 = b == kXBuckets - 1 ? cnt : min(max((int)tab[b + 1], s0), cnt);   // clamped: the table
                }                                                   // is undefined while a chunk is re-run larger
                while (__any_sync(FULL, searcher && s1 - s0 > 4)) {
                    const bool act = searcher && s1 - s0 > 4;
                    const int mid = (s0 + s1) >> 1;
                    float x = 0.f;
                    if (act) x = __ldg(&A.cand[coff + mid].x);
                    const bool below = which == 0 ? (x < key) : (x <= key);
                    if (act) { if (below) s0 = mid + 1; else s1 = mid; }
                }
                if (G == 1) { if (narrow) { if (pass == 0) lo = s0; else hi = s1; } }
                else {
                    const int l0 = __shfl_sync(FULL, s0, lane - gl);
                    const int u1 = __shfl_sync(FULL, s1, lane - gl + 1);
                    if (narrow) { lo = l0; hi = u1; }
                }
            }
        }
        // ---- chunk numbering: exclusive prefix sum of the chunk counts over the batch ----------
        const int len = boxed ? max(hi - lo, 0) : 0;
        const int nch = (len + kChunk - 1) / kChunk;
        const int mine = gl == 0 ? nch : 0;
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const int total = __shfl_sync(FULL, inc, 31);
        const bool mapped = total <= kOwnerCap;            // warp uniform
        if (gl == 0) {
            if (mapped)
                for (int c = 0; c < nch; c++) W.owner[inc - mine + c] = (unsigned char)grp;
            W.cstart[grp] = inc - mine;
            W.clen[grp] = len;
            W.cbase[grp] = coff + lo;
            W.blo[grp] = make_float4(bx0, by0, bz0, 0.f);
            W.bhi[grp] = make_float4(bx1, by1, bz1, 0.f);
        }
        __syncwarp();
        // ---- load-balanced scan: lane <- chunk, whatever partial match owns it ------------------
        for (int cb = 0; cb < total; cb += 32 / kQuad) {
            const int ch = cb + (lane >> 2), sub = lane & (kQuad - 1);
            const bool act = ch < total;
            int g = 0;
            if (mapped) {
                g = act ? (int)W.owner[ch] : 0;
            } else {
#pragma unroll
                for (int s = IPW / 2; s > 0; s >>= 1)
                    if (W.cstart[g + s] <= ch) g += s;
            }
            const int k0 = (ch - W.cstart[g]) * kChunk;
            const int rem = act ? W.clen[g] - k0 : 0;          // records left from the chunk start
            const float4 *src = A.cand + W.cbase[g] + k0;
            constexpr int NR = kChunk / kQuad;                 // records per lane
            float4 r[NR];
#pragma unroll
            for (int i = 0; i < NR; i++) {
                r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sub + kQuad * i < rem) r[i] = __ldg(src + sub + kQuad * i);
            }
            const float4 b0 = W.blo[g], b1 = W.bhi[g];
            bool ok[NR];
            unsigned m[NR], many = 0;
#pragma unroll
            for (int i = 0; i < NR; i++) {
                ok[i] = sub + kQuad * i < rem && r[i].x >= b0.x && r[i].x <= b1.x && r[i].y >= b0.y && r[i].y <= b1.y &&
                        r[i].z >= b0.z && r[i].z <= b1.z;
                m[i] = __ballot_sync(FULL, ok[i]);
                many |= m[i];
            }
            if (stats) {
                int tsum = 0;
#pragma unroll
                for (int i = 0; i < NR; i++) tsum += (sub + kQuad * i < rem) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(FULL, tsum, o);
                if (lane == 0) {
                    n_scan += (unsigned)tsum;
#pragma unroll
                    for (int i = 0; i < NR; i++) n_box += __popc(m[i]);
                }
            }
            if (many) {
                int pos = qfill;
#pragma unroll
                for (int i = 0; i < NR; i++) {
                    if (ok[i]) W.queue[pos + __popc(m[i] & lt)] = make_int2(__float_as_int(r[i].w), g);
                    pos += __popc(m[i]);
                }
                qfill = pos;
                while (qfill >= 32) { qfill -= 32; drain(qfill, 32); }
            }
        }
        if (qfill > 0) { drain(0, qfill); qfill = 0; }
        __syncwarp();   // the next batch overwrites the per-warp tables
        if (stats && lane == 0 && (n_scan | n_box)) {   // unsigned partials: publish per batch
            atomicAdd(&A.ctr->scan_tests, (u64)n_scan);
            atomicAdd(&A.ctr->box_pass, (u64)n_box);
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
