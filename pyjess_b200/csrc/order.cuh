// order.cuh -- the reference's ITERATION ORDER of the tuples of one (structure, template)
// pair, reconstructed on the device (jb_query_params.hit_order = JB_ORDER_REFERENCE*).
//
// The breadth-first engine finds the same SET of tuples as the reference's depth-first
// Scanner_next (eff:Scanner.c:193-304), but two reference features depend on the ORDER in
// which that scanner yields them: `max_candidates` (src/pyjess/_jess.pyx:2263-2269) and the
// tie rule of `best_match` (:2238, first of equal RMSDs wins) -- and a drop-in `Query`
// should iterate like the original.  The depth-first order is a lexicographic order:
//
//   level 0      candidates in CandidateSet order (eff:Scanner.c:212-227): the order in which
//                TessTemplate_candidates appended them (eff:TessTemplate.c:141-201) -- for
//                match modes -1..8 the template atom's residue names in the order given
//                (duplicates skipped), inside one residue name by (serial, atom index)
//                (ResIndex_create's sort, eff:ResIndex.c:18-27,58-61); other modes: atom order;
//   level l > 0  the order in which KdTreeQuery_next pops leaves (eff:KdTree.c:238-296): it
//                pushes left then right and pops from the top, i.e. right subtree first, so
//                the leaves come in REVERSE left-to-right order of the tree that
//                KdTreeNode_create built over the candidate set (eff:KdTree.c:343-423):
//                median split by qselect_r (patches/qselect.h.patch:60-119: median-of-three
//                pivot, Lomuto partition with `<=`), axis cycling x,y,z with the depth.
//                Box pruning only skips subtrees, it never reorders.
//
// So every candidate of a list owns two fixed ranks -- its CandidateSet position and the
// reverse leaf position of the k-d tree -- and the tuples of a pair sort by the rank vector
// taken in the reference's scan order (eff:Scanner.c:88-99,160-174), whatever scan order the
// engine itself used.  Lists are keyed by (structure, ORDERED spec): the engine's candidate
// lists are keyed by a spec with its residue names as a SET, the reference's list order also
// depends on the order the names were written in.
//
//   k_olist_count / scan   size and offset of every (structure, ordered spec) list
//   k_olist_fill           records (list, atom) + CandidateSet sort key
//   cub radix sort x2      (list, atom) ascending -> lookup table; then stable by
//                          (list, residue-name position, serial) -> CandidateSet order
//   k_olist_ranks          CandidateSet position of every record + FP64 coordinates in that order
//   k_kd_order             one thread per list replays KdTreeNode_create's index permutation
//   k_hit_keys             rank vector of every hit; hits.cuh then sorts by it
#pragma once
#include "kernels.cuh"

namespace jb {

struct OrderSpecs {            // ordered specs of a template set (device)
    int n_ospecs;
    const int *spec;           // [OS] engine spec id (candidate list to read)
    const int *by_res;         // [OS] 1: match mode -1..8 (list built through the residue index)
    const int *res_off;        // [OS+1]
    const uint32_t *res_up;    // residue names in the order written, upper-cased, duplicates removed
    const int *atom_ospec;     // [TA] ordered spec of every template atom
};

constexpr int kListBits = 27;  // lists per chunk < 2^27 (checked on the host)

__global__ void __launch_bounds__(256) k_iota(unsigned *v, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (unsigned)i;
}

// The rank tables are only read for hits that SHARE their (structure, template) pair with another hit -- alone in its
// pair a hit needs no order.  k_pair_hits counts the hits of every pair, k_mark_lists flags the lists of the pairs with
// two or more; the lists nobody asks for stay empty (sparse results -- BASELINE config 2: 97 hits, a 58,870-atom
// structure -- spent 10 of their 14 ms replaying k-d trees whose ranks were never read).
__global__ void __launch_bounds__(256) k_pair_hits(long long h0, long long h1, const int *pair, unsigned *pair_hits)
{
    const long long h = h0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (h < h1) atomicAdd(&pair_hits[pair[h]], 1u);
}

__global__ void __launch_bounds__(256) k_mark_lists(long long h0, long long h1, const int *pair, const unsigned *pair_hits,
                                                    const PairRec *recs, OrderSpecs O, unsigned char *needed)
{
    const long long h = h0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= h1 || pair_hits[pair[h]] < 2u) return;
    const PairRec *rec = recs + pair[h];
    for (int k = 0; k < rec->n; k++) needed[(long long)rec->m * O.n_ospecs + O.atom_ospec[rec->ta0 + k]] = 1;
}

__global__ void __launch_bounds__(256) k_olist_count(int Mc, OrderSpecs O, int n_specs, const CandHead *head,
                                                     const unsigned char *needed, unsigned *ocnt)
{
    const long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nl = (long long)Mc * O.n_ospecs;
    if (L > nl) return;
    unsigned c = 0;
    if (L < nl && needed[L]) {
        const int m = (int)(L / O.n_ospecs), os = (int)(L % O.n_ospecs);
        c = (unsigned)head[(long long)m * n_specs + O.spec[os]].cnt;
    }
    ocnt[L] = c;               // entry nl = 0: becomes the total after the exclusive scan
}

// warp per list: one record per candidate
__global__ void __launch_bounds__(256) k_olist_fill(int Mc, OrderSpecs O, int n_specs, const CandHead *head,
                                                    const float4 *cand, const long long *atom_off,
                                                    const uint32_t *res, const int *serial, const unsigned *ooff,
                                                    const unsigned char *needed, u64 *key_atom, u64 *key_set)
{
    const long long L = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (L >= (long long)Mc * O.n_ospecs || !needed[L]) return;
    const int m = (int)(L / O.n_ospecs), os = (int)(L % O.n_ospecs);
    const CandHead hd = head[(long long)m * n_specs + O.spec[os]];
    const long long a0 = atom_off[m];
    const int r0 = O.res_off[os], r1 = O.res_off[os + 1];
    const bool by_res = O.by_res[os] != 0;
    const unsigned base = ooff[L];
    for (int j = lane; j < hd.cnt; j += 32) {
        const int a = __float_as_int(cand[hd.off + j].w);
        u64 minor;
        if (by_res) {
            const uint32_t tok = upper4(res[a0 + a]);
            int rp = 31;
            for (int r = r0; r < r1; r++)
                if (O.res_up[r] == tok) { rp = r - r0; break; }
            const int ser = serial ? serial[a0 + a] : a + 1;
            minor = ((u64)rp << 32) | (u64)((uint32_t)ser ^ 0x80000000u);
        } else {
            minor = (u64)(uint32_t)a;                      // molecule order
        }
        key_atom[base + j] = ((u64)L << 32) | (u64)(uint32_t)a;
        key_set[base + j] = ((u64)L << 37) | minor;
    }
}

// after the two sorts: position j of the CandidateSet order holds record perm[j] of the
// (list, atom) table.  Writes that record's CandidateSet position and stages the FP64
// coordinates of the list in CandidateSet order (CandidateSet_recordCoordinates).
__global__ void __launch_bounds__(256) k_olist_ranks(long long n_rec, int n_ospecs, const u64 *set_sorted,
                                                     const unsigned *perm, const u64 *atom_sorted, const unsigned *ooff,
                                                     const long long *atom_off, const AtomRec *atoms, unsigned *set_pos,
                                                     double *coords, int *kd_idx)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_rec) return;
    const long long L = (long long)(set_sorted[j] >> 37);
    const unsigned i = perm[j];
    const unsigned local = (unsigned)(j - ooff[L]);
    set_pos[i] = local;
    kd_idx[j] = (int)local;
    const int m = (int)(L / n_ospecs);
    const AtomRec ar = atoms[atom_off[m] + (long long)(uint32_t)atom_sorted[i]];
    coords[3 * j] = ar.x; coords[3 * j + 1] = ar.y; coords[3 * j + 2] = ar.z;
}

// ---- qselect_r / KdTreeNode_create, index permutation only -------------------------------
// compare(a, b) of KdTree_compare_r (eff:KdTree.c:312-321): sign of u[a][axis] - u[b][axis]
__device__ __forceinline__ int kd_cmp(const double *u, int a, int b, int axis)
{
    const double c = u[3 * a + axis], d = u[3 * b + axis];
    return c < d ? -1 : c > d ? 1 : 0;
}

__device__ __forceinline__ void kd_swap(int *idx, int i, int j)
{
    const int t = idx[i]; idx[i] = idx[j]; idx[j] = t;
}

// patches/qselect.h.patch:99-119 with :60-97 inlined; returns nothing: the caller only needs
// the permutation (the returned index is always k, see KdTreeNode_create below)
__device__ void kd_qselect(int *idx, int n, int k, const double *u, int axis)
{
    int left = 0, right = n - 1;
    while (left != right) {
        // _qselect_median3
        const int mid = (left + right + 1) / 2;
        int pivot;
        if ((kd_cmp(u, idx[left], idx[mid], axis) > 0) != (kd_cmp(u, idx[left], idx[right], axis) > 0)) pivot = left;
        else if ((kd_cmp(u, idx[mid], idx[left], axis) < 0) != (kd_cmp(u, idx[mid], idx[right], axis) < 0)) pivot = mid;
        else pivot = right;
        // _qselect_partition
        kd_swap(idx, right, pivot);
        int store = left;
        const double pv = u[3 * idx[right] + axis];
        for (int i = left; i < right; i++) {
            const double c = u[3 * idx[i] + axis];
            if (!(c > pv)) {                               // compare(i, right) <= 0
                kd_swap(idx, i, store);
                store++;
            }
        }
        kd_swap(idx, store, right);
        if (store == k) return;
        if (k < store) right = store - 1; else left = store + 1;
    }
}

// One thread per list.  KdTreeNode_create (eff:KdTree.c:343-423): n == 1 leaf; otherwise
// qselect the median position n/2 on the current axis, recurse on [0, n/2) and [n/2, n) with
// the next axis.  The final index array is the left-to-right leaf order; KdTreeQuery_next
// visits leaves right to left, so rank = count - 1 - leaf position.
__global__ void __launch_bounds__(128) k_kd_order(long long n_lists, const unsigned *schedule, const unsigned *ooff,
                                                  const double *coords, int *kd_idx, const unsigned *perm, unsigned *kd_rank,
                                                  unsigned skip_lo, unsigned skip_hi)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_lists) return;
    const long long L = schedule ? (long long)schedule[t] : t;
    const unsigned o0 = ooff[L], cnt = ooff[L + 1] - o0;
    if (cnt == 0 || (cnt > skip_lo && cnt <= skip_hi)) return;   // (skip_lo, skip_hi]: lists of the team kernels
    int *idx = kd_idx + o0;
    const double *u = coords + 3ll * o0;
    int st_s[40], st_n[40];
    unsigned char st_d[40];
    int sp = 0;
    st_s[0] = 0; st_n[0] = (int)cnt; st_d[0] = 0; sp = 1;
    while (sp > 0) {
        sp--;
        const int s = st_s[sp], n = st_n[sp], d = st_d[sp];
        if (n <= 1) continue;
        const int split = n / 2;
        kd_qselect(idx + s, n, split, u, d % 3);
        // larger half first on the stack keeps it shallow; the halves are independent
        st_s[sp] = s + split; st_n[sp] = n - split; st_d[sp] = (unsigned char)(d + 1); sp++;
        st_s[sp] = s; st_n[sp] = split; st_d[sp] = (unsigned char)(d + 1); sp++;
    }
    for (unsigned q = 0; q < cnt; q++) kd_rank[perm[o0 + idx[q]]] = cnt - 1 - q;
}

// One WARP per list of kd_team_lo < cnt <= CAP records.  A thread walking a 1,000-record list alone (n log n dependent
// gathers from global memory) was the whole cost of the thread-per-list kernel: 3.6 ms per launch on BASELINE config 3,
// 100 ms on config 4's large structures (profiles/r2_launches.md).  Here the index array lives in shared memory and the
// tree is built LEVEL BY LEVEL: the segments of one level are disjoint, so lane t replays qselect on segment t -- the same
// comparisons on the same FP64 values in the same order as KdTreeNode_create, whatever the order the segments are taken
// in.  The top levels (1, 2, 4 ... segments) still run on few lanes, but against shared memory and the L1-resident
// coordinates of ONE list instead of 128 lists per CTA.
constexpr int kKdThreadMax = 32;                 // lists up to here stay with the thread-per-list kernel

__device__ __forceinline__ void kd_qselect_sm(int *idx, int n, int k, const double *u, int axis)
{
    int left = 0, right = n - 1;
    while (left != right) {
        const int mid = (left + right + 1) / 2;
        int pivot;
        if ((kd_cmp(u, idx[left], idx[mid], axis) > 0) != (kd_cmp(u, idx[left], idx[right], axis) > 0)) pivot = left;
        else if ((kd_cmp(u, idx[mid], idx[left], axis) < 0) != (kd_cmp(u, idx[mid], idx[right], axis) < 0)) pivot = mid;
        else pivot = right;
        kd_swap(idx, right, pivot);
        int store = left;
        const double pv = __ldg(&u[3 * idx[right] + axis]);
        // the elements of steps i + 1 and i + 2 are never touched by the swap of step i (store <= i): fetch them ahead
        int v = idx[left], v1 = idx[min(left + 1, right)];
        double c = __ldg(&u[3 * v + axis]), c1 = __ldg(&u[3 * v1 + axis]);
        for (int i = left; i < right; i++) {
            const int v2 = idx[min(i + 2, right)];                           // (>= right: the pivot, unused)
            const double c2 = __ldg(&u[3 * v2 + axis]);
            if (!(c > pv)) {
                idx[i] = idx[store];
                idx[store] = v;
                store++;
            }
            v = v1; c = c1; v1 = v2; c1 = c2;
        }
        kd_swap(idx, store, right);
        if (store == k) return;
        if (k < store) right = store - 1; else left = store + 1;
    }
}

template <int CAP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_kd_order_team(long long n_lists, int lo, int hi, const unsigned *ooff,
                                                              const double *coords, const unsigned *perm, unsigned *kd_rank)
{
    extern __shared__ int kd_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *idx = kd_sm + warp * (2 * CAP);                                     // [CAP]
    static_assert(CAP <= 16384, "segments are packed into 32 bits");
    unsigned *qa = reinterpret_cast<unsigned *>(idx + CAP), *qb = qa + CAP / 2;   // segment queues of the current / next level
    const unsigned lt = (1u << lane) - 1;
    for (long long L = (long long)blockIdx.x * WARPS + warp; L < n_lists; L += (long long)gridDim.x * WARPS) {
        const unsigned o0 = ooff[L], cnt = ooff[L + 1] - o0;
        if ((int)cnt <= lo || (int)cnt > hi || cnt > (unsigned)CAP) continue;   // warp uniform: lists of (lo, hi] are this launch's
        const double *u = coords + 3ll * o0;
        for (int j = lane; j < (int)cnt; j += 32) idx[j] = j;
        // a segment is (start << 16) | n with n >= 2 (CAP <= 16,384)
        if (lane == 0) qa[0] = cnt;
        int n_cur = 1;
        __syncwarp();
        for (int depth = 0; n_cur > 0; depth++) {
            int n_next = 0;
            for (int i0 = 0; i0 < n_cur; i0 += 32) {
                const int i = i0 + lane;
                unsigned left = 0, right = 0;                                 // children worth another split
                if (i < n_cur) {
                    const unsigned seg = qa[i];
                    const int s = (int)(seg >> 16), n = (int)(seg & 0xffffu);
                    const int split = n / 2;
                    kd_qselect_sm(idx + s, n, split, u, depth % 3);
                    if (split >= 2) left = ((unsigned)s << 16) | (unsigned)split;
                    if (n - split >= 2) right = ((unsigned)(s + split) << 16) | (unsigned)(n - split);
                }
                const unsigned bl = __ballot_sync(0xffffffffu, left != 0), br = __ballot_sync(0xffffffffu, right != 0);
                if (left) qb[n_next + __popc(bl & lt)] = left;
                n_next += __popc(bl);
                if (right) qb[n_next + __popc(br & lt)] = right;
                n_next += __popc(br);
            }
            __syncwarp();
            unsigned *t = qa; qa = qb; qb = t;
            n_cur = n_next;
        }
        for (int q = lane; q < (int)cnt; q += 32) kd_rank[perm[o0 + idx[q]]] = cnt - 1 - q;
        __syncwarp();
    }
}

// Small sparse results: does any (structure, template) pair own more than one hit?  If not, the
// record order inside a pair is moot and the rank tables need not be built (a single small query
// spends more time in them than in the search).  One CTA, n <= 1024 hits.
__global__ void __launch_bounds__(256) k_hit_multi(long long h0, int n, const int *mol, const int *tpl, int *multi)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int m = mol[h0 + i], t = tpl[h0 + i];
        for (int j = i + 1; j < n; j++)
            if (mol[h0 + j] == m && tpl[h0 + j] == t) { *multi = 1; return; }
    }
}

struct HitKeyArgs {
    long long h0, h1;
    int stride, n_specs, ref_reorder;      // ref_reorder: the reference ran with reorder=True
    const int *pair, *atoms;               // unordered hit arrays (k_superpose)
    const PairRec *recs;
    const int *tpl_spec;                   // TemplateTable.spec
    const CandHead *head;
    OrderSpecs O;
    const unsigned *ooff;
    const u64 *atom_sorted;
    const unsigned *set_pos, *kd_rank;
    int *keys;                             // [hit_cap * stride], indexed like atoms
};

// thread per hit: rank vector in the REFERENCE's scan order
__global__ void __launch_bounds__(128) k_hit_keys(HitKeyArgs A)
{
    const long long h = A.h0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= A.h1) return;
    const PairRec *rec = A.recs + A.pair[h];
    const int m = rec->m, n = rec->n, ta0 = rec->ta0;
    int cnt[kMaxN];
    unsigned char ro[kMaxN];
    for (int k = 0; k < n; k++) {
        cnt[k] = A.head[(long long)m * A.n_specs + A.tpl_spec[ta0 + k]].cnt;
        ro[k] = (unsigned char)k;
    }
    if (A.ref_reorder) {                   // argsort by (|C_k|, k): eff:Scanner.c:88-99
        for (int i = 1; i < n; i++) {
            const int v = ro[i];
            int j = i - 1;
            while (j >= 0 && (cnt[ro[j]] > cnt[v] || (cnt[ro[j]] == cnt[v] && ro[j] > v))) { ro[j + 1] = ro[j]; j--; }
            ro[j + 1] = (unsigned char)v;
        }
    }
    int *key = A.keys + h * A.stride;
    const int *at = A.atoms + h * A.stride;
    for (int l = 0; l < A.stride; l++) {
        int v = 0;
        if (l < n) {
            const int k = ro[l];
            const long long L = (long long)m * A.O.n_ospecs + A.O.atom_ospec[ta0 + k];
            const u64 want = ((u64)L << 32) | (u64)(uint32_t)at[k];
            const unsigned end = A.ooff[L + 1];
            unsigned lo = A.ooff[L], hi = end;
            while (lo < hi) {
                const unsigned mid = (lo + hi) >> 1;
                if (A.atom_sorted[mid] < want) lo = mid + 1; else hi = mid;
            }
            // (a list that was not built belongs to a pair with this single hit: any key will do)
            if (lo < end) v = (int)(l == 0 ? A.set_pos[lo] : A.kd_rank[lo]);
        }
        key[l] = v;
    }
}

}  // namespace jb
