// hits.cuh -- canonical ordering of the compacted hits ON THE DEVICE.
//
// k_superpose (kernels.cuh) appends hits in the arbitrary order its warps finish.
// The C ABI promises (molecule, template, atom tuple) ascending order
// (include/jess_b200.h: jb_hits), which is also what makes the result comparable
// with the reference's hit set "after canonical ordering" (BASELINE north_star).
// Sorting 10^5..10^7 records of ~200 bytes on the host dominated the end-to-end time,
// so the order is established here with a counting sort over the (structure,
// template) pairs of the chunk followed by a rank inside each -- usually tiny -- run:
//
//   k_hit_key     64-bit key (pair, first key column, second key column) per hit
//   radix sort    cub::DeviceRadixSort over the keys: hits grouped by pair AND ordered by the first
//                 two columns in one pass over the data (dense-hit workloads used to spend more
//                 time in the segmented bitonic sort of whole pairs than in the search itself)
//   k_run_heads   runs of EQUAL keys (ties on the first two columns) -> run ids and offsets (scan)
//   k_hit_rank    ties: rank of every hit inside its (usually 1-element) run by full tuple compare
//   k_run_* / k_tile_*  long tie runs: segmented, tiled bitonic sort of the index lists
//   k_hit_gather  hit records -> final arrays in canonical order (coalesced writes)
// (k_hit_hist / k_hit_group: the counting-sort grouping by pair, kept for structures whose atom
//  indices do not fit the key's 20/21-bit columns.)
#pragma once
#include "kernels.cuh"

namespace jb {

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanPerThread;   // 2048 elements per CTA

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned *total)
{
    __shared__ unsigned warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < kScanThreads / 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const unsigned base = w > 0 ? warp_sums[w - 1] : 0;
    *total = warp_sums[kScanThreads / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// phase 1: per-tile exclusive scan in place + tile totals
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(unsigned *data, long long n, unsigned *tile_total)
{
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanPerThread;
    unsigned v[kScanPerThread], sum = 0;
#pragma unroll
    for (int i = 0; i < kScanPerThread; i++) {
        v[i] = base + i < n ? data[base + i] : 0u;
        sum += v[i];
    }
    unsigned total;
    unsigned run = block_exclusive_scan(sum, &total);
#pragma unroll
    for (int i = 0; i < kScanPerThread; i++) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0) tile_total[blockIdx.x] = total;
}

// phase 2: one CTA turns the tile totals into tile offsets (sequential over 256-wide slabs)
__global__ void __launch_bounds__(kScanThreads) k_scan_totals(unsigned *tile_total, int n_tiles)
{
    unsigned carry = 0;
    for (int s = 0; s < n_tiles; s += kScanThreads) {
        const int i = s + threadIdx.x;
        const unsigned v = i < n_tiles ? tile_total[i] : 0u;
        unsigned total;
        const unsigned ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_total[i] = carry + ex;
        carry += total;
    }
}

// phase 3: add the tile offsets
__global__ void __launch_bounds__(kScanThreads) k_scan_add(unsigned *data, long long n, const unsigned *tile_total)
{
    const unsigned add = tile_total[blockIdx.x];
    const long long base = (long long)blockIdx.x * kScanTile;
    for (int i = threadIdx.x; i < kScanTile; i += kScanThreads)
        if (base + i < n) data[base + i] += add;
}

struct HitOrderArgs {
    long long h0, h1;           // hit range of the chunk in the unordered arrays
    int mol_base;               // first structure of the chunk
    int n_templates;
    int stride;
    const int *mol, *tpl, *atoms;
    unsigned *off;              // [pairs + 1] counts, then exclusive offsets
    unsigned *fill;             // [pairs] tickets
    unsigned *grouped;          // [h1 - h0] hit indices grouped by run
    unsigned *order;            // [h1 - h0] hit indices in canonical order
    const unsigned *runid;      // [h1 - h0] run of every GROUPED position (NULL: runs = pairs, off indexed by pair)
};

__device__ __forceinline__ long long hit_pair(const HitOrderArgs &A, long long i)
{
    return (long long)(A.mol[i] - A.mol_base) * A.n_templates + A.tpl[i];
}

__global__ void __launch_bounds__(256) k_hit_hist(HitOrderArgs A)
{
    const long long i = A.h0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.h1) atomicAdd(&A.off[hit_pair(A, i)], 1u);
}

__global__ void __launch_bounds__(256) k_hit_group(HitOrderArgs A)
{
    const long long i = A.h0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.h1) return;
    const long long p = hit_pair(A, i);
    A.grouped[A.off[p] + atomicAdd(&A.fill[p], 1u)] = (unsigned)(i - A.h0);
}

// ---- keyed path: (pair | column 0 | column 1) ------------------------------------------------
constexpr int kKeyCol0Bits = 20, kKeyCol1Bits = 21;      // 23 bits remain for the pair of the chunk

__global__ void __launch_bounds__(256) k_hit_key(HitOrderArgs A, u64 *key, unsigned *iota)
{
    const long long i = A.h0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.h1) return;
    const int *row = A.atoms + i * A.stride;
    const u64 c0 = (u64)((unsigned)row[0] & ((1u << kKeyCol0Bits) - 1));
    const u64 c1 = A.stride > 1 ? (u64)((unsigned)row[1] & ((1u << kKeyCol1Bits) - 1)) : 0ull;
    key[i - A.h0] = ((u64)hit_pair(A, i) << (kKeyCol0Bits + kKeyCol1Bits)) | (c0 << kKeyCol1Bits) | c1;
    iota[i - A.h0] = (unsigned)(i - A.h0);
}

// Dense-hit fallback of the keyed path: when the first sort leaves LONG tie runs (thousands of hits
// of one pair sharing their first two columns: symmetric side chains of large templates), the
// whole chunk is ordered by a multi-pass LSD radix sort instead -- trailing columns first, `per`
// columns of `bits` bits per 64-bit key, every pass stable, the (pair, c0, c1) pass last.
__global__ void __launch_bounds__(256) k_lsd_key(HitOrderArgs A, const unsigned *perm, int lo, int hi, int bits, int final,
                                                 u64 *key)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.h1 - A.h0) return;
    const long long i = A.h0 + perm[j];
    const int *row = A.atoms + i * A.stride;
    u64 k = 0;
    if (final) {
        const u64 c0 = (u64)((unsigned)row[0] & ((1u << kKeyCol0Bits) - 1));
        const u64 c1 = A.stride > 1 ? (u64)((unsigned)row[1] & ((1u << kKeyCol1Bits) - 1)) : 0ull;
        k = ((u64)hit_pair(A, i) << (kKeyCol0Bits + kKeyCol1Bits)) | (c0 << kKeyCol1Bits) | c1;
    } else {
        const unsigned mask = (1u << bits) - 1;
        for (int c = lo; c < hi; c++) k = (k << bits) | (u64)((unsigned)row[c] & mask);
    }
    key[j] = k;
}

// flag[j] = 1 where a new run of equal keys starts (j > 0); scanned in place afterwards
__global__ void __launch_bounds__(256) k_run_flags(long long n, const u64 *skey, unsigned *flag)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n) return;
    flag[j] = (j > 0 && j < n && skey[j] != skey[j - 1]) ? 1u : 0u;
}

// after the exclusive scan of the flags: run id of every position, run offsets, run count
__global__ void __launch_bounds__(256) k_run_heads(long long n, const u64 *skey, const unsigned *scanned, unsigned *runid,
                                                   unsigned *off, unsigned *n_runs)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const bool head = j == 0 || skey[j] != skey[j - 1];
    const unsigned r = scanned[j] + ((j > 0 && head) ? 1u : 0u);
    runid[j] = r;
    if (head) off[r] = (unsigned)j;
    if (j == n - 1) { off[r + 1] = (unsigned)n; *n_runs = r + 1; }
}

constexpr unsigned kShortRun = 48;   // runs up to this length are ranked, longer ones sorted by a CTA

// other < mine in lexicographic tuple order
__device__ __forceinline__ bool tuple_less(const int *other, const int *mine, int stride)
{
    for (int k = 0; k < stride; k++) {
        const int a = other[k], b = mine[k];
        if (a != b) return a < b;
    }
    return false;
}

__global__ void __launch_bounds__(256) k_hit_rank(HitOrderArgs A)
{
    const long long q = A.h0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= A.h1) return;
    // keyed path: one thread per GROUPED position; counting-sort path: one thread per hit
    const long long i = A.runid ? A.h0 + A.grouped[q - A.h0] : q;
    const long long p = A.runid ? (long long)A.runid[q - A.h0] : hit_pair(A, i);
    const unsigned r0 = A.off[p], r1 = A.off[p + 1];
    if (r1 - r0 == 1) { A.order[r0] = (unsigned)(i - A.h0); return; }
    if (r1 - r0 > kShortRun) return;                          // the tiled sort owns this run
    const int *mine = A.atoms + i * A.stride;
    unsigned rank = 0;
    for (unsigned r = r0; r < r1; r++) {
        const long long j = A.h0 + A.grouped[r];
        if (j != i && tuple_less(A.atoms + j * A.stride, mine, A.stride)) rank++;
    }
    A.order[r0 + rank] = (unsigned)(i - A.h0);
}

// ---- long runs (dense-hit pairs: 10^2 .. 10^5 hits of ONE template in ONE structure) ----
// Segmented bitonic sort of the index lists, tiled: a run of n indices is cut into tiles
// of kTile; k_run_tiles builds the (run, tile) work list; k_tile_sort sorts every tile in
// shared memory (all stages with k <= kTile); each larger merge k runs its wide stages
// (j >= kTile) as one grid-wide launch per stage (k_run_stage) and its narrow stages
// (j < kTile) again inside shared memory (k_tile_merge).  The network is the all-ascending
// variant (first stage of a merge pairs i with i ^ (k-1), the rest i ^ j), so positions
// beyond n act as +infinity and never move: no physical padding.  Comparisons look at a
// 64-bit prefix (first two atoms) cached in shared memory first and gather the full tuples
// from L2 only on a tie.
constexpr int kTile = 2048;
constexpr int kTileThreads = 512;

struct alignas(16) RunTile {
    unsigned r0;     // start of the run in grouped[]
    unsigned n;      // run length
    unsigned tile;   // tile index inside the run
    unsigned npad;   // n rounded up to a power of two
};

struct RunWork {
    unsigned n_tiles;
    unsigned max_npad;
    unsigned long_hits;      // hits inside runs longer than kShortRun
    unsigned n_runs;         // keyed path: number of tie runs (written by k_run_heads)
};

__global__ void __launch_bounds__(256) k_run_tiles(HitOrderArgs A, long long n_pairs, const unsigned *n_runs, RunTile *tiles,
                                                   unsigned cap, RunWork *work)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_runs) n_pairs = (long long)*n_runs;
    if (p >= n_pairs) return;
    const unsigned r0 = A.off[p], n = A.off[p + 1] - r0;
    if (n <= kShortRun) return;
    unsigned npad = 1;
    while (npad < n) npad <<= 1;
    const unsigned nt = (n + kTile - 1) / kTile;
    const unsigned base = atomicAdd(&work->n_tiles, nt);
    atomicMax(&work->max_npad, npad);
    atomicAdd(&work->long_hits, n);
    for (unsigned t = 0; t < nt; t++)
        if (base + t < cap) { RunTile e; e.r0 = r0; e.n = n; e.tile = t; e.npad = npad; tiles[base + t] = e; }
}

__device__ __forceinline__ u64 tuple_prefix(const int *row, int stride)
{
    const unsigned a = (unsigned)row[0] ^ 0x80000000u;                      // signed order -> unsigned order
    const unsigned b = stride > 1 ? ((unsigned)row[1] ^ 0x80000000u) : 0u;
    return ((u64)a << 32) | b;
}

// x > y for hits x, y (indices relative to h0) with cached prefixes
__device__ __forceinline__ bool hit_greater(const HitOrderArgs &A, unsigned x, u64 kx, unsigned y, u64 ky)
{
    if (kx != ky) return kx > ky;
    return tuple_less(A.atoms + (A.h0 + y) * A.stride, A.atoms + (A.h0 + x) * A.stride, A.stride);
}

// stages of one tile in shared memory: merges k_lo..k_hi, each from stage j_hi(k) down to 1
template <bool FULL_SORT>
__device__ __forceinline__ void tile_stages(const HitOrderArgs &A, const RunTile e, unsigned k_merge, unsigned *s_idx,
                                            u64 *s_key)
{
    const unsigned base = e.tile * kTile;
    const unsigned count = min((unsigned)kTile, e.n - base);
    unsigned *idx = A.grouped + e.r0 + base;
    for (unsigned i = threadIdx.x; i < count; i += kTileThreads) {
        const unsigned h = idx[i];
        s_idx[i] = h;
        s_key[i] = tuple_prefix(A.atoms + (A.h0 + h) * A.stride, A.stride);
    }
    __syncthreads();
    const unsigned k_first = FULL_SORT ? 2u : k_merge, k_last = FULL_SORT ? (unsigned)kTile : k_merge;
    for (unsigned k = k_first; k <= k_last && k <= e.npad; k <<= 1) {
        for (unsigned j = FULL_SORT ? (k >> 1) : (unsigned)(kTile >> 1); j > 0; j >>= 1) {
            const unsigned flip = (j == (k >> 1)) ? (k - 1) : j;             // global index pairing
            for (unsigned i = threadIdx.x; i < (unsigned)kTile; i += kTileThreads) {
                const unsigned gi = base + i, gq = gi ^ flip;                // stays inside the tile: flip < kTile
                const unsigned q = gq - base;
                if (gq > gi && gq < e.n) {
                    const unsigned x = s_idx[i], y = s_idx[q];
                    const u64 kx = s_key[i], ky = s_key[q];
                    if (hit_greater(A, x, kx, y, ky)) { s_idx[i] = y; s_idx[q] = x; s_key[i] = ky; s_key[q] = kx; }
                }
            }
            __syncthreads();
        }
    }
    for (unsigned i = threadIdx.x; i < count; i += kTileThreads) idx[i] = s_idx[i];
}

__global__ void __launch_bounds__(kTileThreads) k_tile_sort(HitOrderArgs A, const RunTile *tiles)
{
    __shared__ unsigned s_idx[kTile];
    __shared__ u64 s_key[kTile];
    tile_stages<true>(A, tiles[blockIdx.x], 0, s_idx, s_key);
}

// narrow stages (j < kTile) of merge k > kTile
__global__ void __launch_bounds__(kTileThreads) k_tile_merge(HitOrderArgs A, const RunTile *tiles, unsigned k)
{
    __shared__ unsigned s_idx[kTile];
    __shared__ u64 s_key[kTile];
    const RunTile e = tiles[blockIdx.x];
    if (k > e.npad) return;
    tile_stages<false>(A, e, k, s_idx, s_key);
}

// one wide stage (j >= kTile) of merge k, all runs at once
__global__ void __launch_bounds__(kTileThreads) k_run_stage(HitOrderArgs A, const RunTile *tiles, unsigned k, unsigned j)
{
    const RunTile e = tiles[blockIdx.x];
    if (k > e.npad) return;
    const unsigned flip = (j == (k >> 1)) ? (k - 1) : j;
    unsigned *idx = A.grouped + e.r0;
    for (unsigned t = threadIdx.x; t < (unsigned)kTile; t += kTileThreads) {
        const unsigned i = e.tile * kTile + t, q = i ^ flip;
        if (q > i && q < e.n) {
            const unsigned x = idx[i], y = idx[q];
            if (tuple_less(A.atoms + (A.h0 + y) * A.stride, A.atoms + (A.h0 + x) * A.stride, A.stride)) {
                idx[i] = y; idx[q] = x;
            }
        }
    }
}

__global__ void __launch_bounds__(kTileThreads) k_run_publish(HitOrderArgs A, const RunTile *tiles)
{
    const RunTile e = tiles[blockIdx.x];
    for (unsigned t = threadIdx.x; t < (unsigned)kTile; t += kTileThreads) {
        const unsigned i = e.tile * kTile + t;
        if (i < e.n) A.order[e.r0 + i] = A.grouped[e.r0 + i];
    }
}

// One warp per output record: lanes copy the 2 + stride ints and 16 doubles of hit
// order[q] to position q of the final arrays.
__global__ void __launch_bounds__(256) k_hit_gather(long long h0, long long h1, int stride, const unsigned *order,
                                                    HitBuf src, HitBuf dst)
{
    const long long q = h0 + ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    if (q >= h1) return;
    const long long s = h0 + order[q - h0];
    if (lane == 0) { dst.mol[q] = src.mol[s]; dst.tpl[q] = src.tpl[s]; dst.rmsd[q] = src.rmsd[s]; }
    for (int k = lane; k < stride; k += 32) dst.atoms[q * stride + k] = src.atoms[s * stride + k];
    if (lane < 9) dst.rot[q * 9 + lane] = src.rot[s * 9 + lane];
    else if (lane < 12) dst.c0[q * 3 + lane - 9] = src.c0[s * 3 + lane - 9];
    else if (lane < 15) dst.c1[q * 3 + lane - 12] = src.c1[s * 3 + lane - 12];
}

// ---- small results: order + gather in ONE launch -------------------------------------------------
// A single-structure query returns a handful of hits; the general ordering path above spends ~10
// launches (+ a device radix sort) on them.  One CTA sorts up to kSmallHits hit indices by (structure,
// template, atom tuple) -- a bitonic network over shared memory, tuples compared in place -- and writes
// the records in that order to `dst`, which may be page-locked HOST memory (the result block itself:
// no device->host copies afterwards).  *multi is set when some (structure, template) holds more than
// one hit: the caller then knows whether the order INSIDE a pair matters (reference hit order).
constexpr int kSmallHits = 1024;

__global__ void __launch_bounds__(256) k_hits_small(long long h0, int nh, int stride, HitBuf src, HitBuf dst, double *dst_xyz_unused,
                                                    int *multi)
{
    __shared__ int s_mol[kSmallHits], s_tpl[kSmallHits], s_idx[kSmallHits];
    (void)dst_xyz_unused;
    int npad = 1;
    while (npad < nh) npad <<= 1;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        s_idx[i] = i < nh ? i : -1;
        s_mol[i] = i < nh ? src.mol[h0 + i] : 0x7fffffff;
        s_tpl[i] = i < nh ? src.tpl[h0 + i] : 0x7fffffff;
    }
    __syncthreads();
    auto less = [&](int a, int b) {                       // hit a before hit b (padding last)
        if (a < 0 || b < 0) return b < 0 && a >= 0;
        if (s_mol[a] != s_mol[b]) return s_mol[a] < s_mol[b];
        if (s_tpl[a] != s_tpl[b]) return s_tpl[a] < s_tpl[b];
        const int *pa = src.atoms + (h0 + a) * stride, *pb = src.atoms + (h0 + b) * stride;
        for (int k = 0; k < stride; k++)
            if (pa[k] != pb[k]) return pa[k] < pb[k];
        return a < b;
    };
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const int x = s_idx[i], y = s_idx[p];
                    const bool up = (i & k) == 0;
                    if (up ? less(y, x) : less(x, y)) { s_idx[i] = y; s_idx[p] = x; }
                }
            }
            __syncthreads();
        }
    for (int q = threadIdx.x; q < nh; q += blockDim.x) {
        const int a = s_idx[q];
        if (q > 0) {
            const int b = s_idx[q - 1];
            if (s_mol[a] == s_mol[b] && s_tpl[a] == s_tpl[b]) *multi = 1;
        }
        const long long s_ = h0 + a, d = h0 + q;
        dst.mol[d] = s_mol[a]; dst.tpl[d] = s_tpl[a]; dst.rmsd[d] = src.rmsd[s_];
        for (int k = 0; k < stride; k++) dst.atoms[d * stride + k] = src.atoms[s_ * stride + k];
        for (int k = 0; k < 9; k++) dst.rot[d * 9 + k] = src.rot[s_ * 9 + k];
        for (int k = 0; k < 3; k++) { dst.c0[d * 3 + k] = src.c0[s_ * 3 + k]; dst.c1[d * 3 + k] = src.c1[s_ * 3 + k]; }
    }
}

// ---- transformed coordinates of the matched atoms -------------------------------
// Hit.atoms(transform=True) (_jess.pyx:2454-2482) builds Hit.transformation = Mv @ Mr @ Mc
// (translation to centre1, rotation, translation by -centre0; :2343-2377) with Mat4.__matmul__
// (:225-240: res[i][j] += a[i][k] * b[k][j], k ascending, from 0) and applies it with
// Atom._transform (:1293-1307: x'[i] = sum_j M[i][j] * (x, 1)[j], j ascending, from 0).  One warp
// per ORDERED hit replays exactly those operations (explicit round-to-nearest, no FMA) on the FP64
// atom records, lane k on the atom matched to template atom k: xyz[(q * stride + k) * 3 ..].
__global__ void __launch_bounds__(256) k_hit_coords(long long h0, long long h1, int stride, int mol_base, MolTable M,
                                                    HitBuf f, double *xyz)
{
    const long long q = h0 + ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    if (q >= h1) return;
    double Mv[4][4], Mr[4][4], Mc[4][4], A[4][4], T[3][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) Mv[i][j] = Mr[i][j] = Mc[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Mc[i][3] = -f.c0[q * 3 + i];
        Mv[i][3] = f.c1[q * 3 + i];
#pragma unroll
        for (int j = 0; j < 3; j++) Mr[i][j] = f.rot[q * 9 + 3 * i + j];
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double r = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) r = __dadd_rn(r, __dmul_rn(Mv[i][k], Mr[k][j]));
            A[i][j] = r;
        }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double r = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) r = __dadd_rn(r, __dmul_rn(A[i][k], Mc[k][j]));
            T[i][j] = r;
        }
    const long long a0 = M.atom_off[f.mol[q] - mol_base];
    for (int k = lane; k < stride; k += 32) {
        const int a = f.atoms[q * stride + k];
        double out[3] = { 0.0, 0.0, 0.0 };
        if (a >= 0) {
            const AtomRec ar = M.atoms[a0 + a];
            const double v[4] = { ar.x, ar.y, ar.z, 1.0 };
#pragma unroll
            for (int i = 0; i < 3; i++) {
                double r = 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++) r = __dadd_rn(r, __dmul_rn(T[i][j], v[j]));
                out[i] = r;
            }
        }
        double *dst = xyz + (q * stride + k) * 3;
        dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2];
    }
}

}  // namespace jb
