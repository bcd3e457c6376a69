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
//   k_hit_hist    count hits per pair                      (atomicAdd, one per hit)
//   scan          exclusive prefix sum over the pairs      (k_scan_*: 3 small kernels)
//   k_hit_group   slot = off[pair] + ticket  -> index list grouped by pair
//   k_hit_rank    rank of every hit inside its run by lexicographic tuple compare
//   k_hit_gather  hit records -> final arrays in canonical order (coalesced writes)
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
    unsigned *grouped;          // [h1 - h0] hit indices grouped by pair
    unsigned *order;            // [h1 - h0] hit indices in canonical order
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

__global__ void __launch_bounds__(256) k_hit_rank(HitOrderArgs A)
{
    const long long i = A.h0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.h1) return;
    const long long p = hit_pair(A, i);
    const unsigned r0 = A.off[p], r1 = A.off[p + 1];
    const int *mine = A.atoms + i * A.stride;
    unsigned rank = 0;
    for (unsigned r = r0; r < r1; r++) {
        const long long j = A.h0 + A.grouped[r];
        if (j == i) continue;
        const int *other = A.atoms + j * A.stride;
        bool less = false;                                    // other < mine, lexicographic
        for (int k = 0; k < A.stride; k++) {
            const int a = other[k], b = mine[k];
            if (a != b) { less = a < b; break; }
        }
        rank += less ? 1u : 0u;
    }
    A.order[r0 + rank] = (unsigned)(i - A.h0);
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

}  // namespace jb
