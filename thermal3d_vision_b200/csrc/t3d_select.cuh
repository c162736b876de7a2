// t3d_select.cuh -- exact order statistics of float32 data by 3-pass radix
// select (11 + 11 + 10 bits of the monotone uint32 key), one CTA per stream.
// Used for np.percentile on non-integer data (utils/preprocessing.py:22) and
// np.median (utils/metrics.py:47, scripts/pseudo_gt.py:174-175).
#pragma once
#include "t3d_common.cuh"

namespace t3d_select {

constexpr int kThreads = 1024;
constexpr int kBins = 2048;

// monotone key: order of keys == numeric order of floats; NaN (either sign bit)
// is handled by the caller (counted separately, sorts last like numpy).
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

struct Smem {
    unsigned int hist[kBins];
    unsigned int warp_tot[kThreads / 32];
    unsigned int sel_bin, sel_rank;
};

// Find bin b with  cum(b-1) <= rank < cum(b); returns (bin, rank - cum(b-1)) in sm.sel_*.
// Block-wide, deterministic. hist has kBins entries; `nb` bins are live.
__device__ __forceinline__ void pick_bin(Smem& sm, unsigned int rank, int nb) {
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    // each thread owns 2 consecutive bins (1024 threads x 2 = 2048)
    const unsigned int h0 = (2 * tid < nb) ? sm.hist[2 * tid] : 0u;
    const unsigned int h1 = (2 * tid + 1 < nb) ? sm.hist[2 * tid + 1] : 0u;
    unsigned int v = h0 + h1, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm.warp_tot[wrp] = incl;
    __syncthreads();
    if (wrp == 0) {
        unsigned int w = sm.warp_tot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        sm.warp_tot[lane] = wi - w;   // exclusive prefix of warp totals
    }
    __syncthreads();
    const unsigned int excl = sm.warp_tot[wrp] + incl - v;   // elements before bin 2*tid
    if (rank >= excl && rank < excl + h0) { sm.sel_bin = 2 * tid; sm.sel_rank = rank - excl; }
    else if (rank >= excl + h0 && rank < excl + v) { sm.sel_bin = 2 * tid + 1; sm.sel_rank = rank - excl - h0; }
    __syncthreads();
}

// pick_bin for a CTA of THREADS threads (a power of two, 64 .. 1024): every thread owns kBins / THREADS consecutive bins.
template <int THREADS>
__device__ __forceinline__ void pick_bin_t(Smem& sm, unsigned int rank, int nb) {
    constexpr int PER = kBins / THREADS;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    unsigned int h[PER], v = 0u;
#pragma unroll
    for (int j = 0; j < PER; ++j) { h[j] = (PER * tid + j < nb) ? sm.hist[PER * tid + j] : 0u; v += h[j]; }
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm.warp_tot[wrp] = incl;
    __syncthreads();
    if (wrp == 0) {
        unsigned int w = (lane < THREADS / 32) ? sm.warp_tot[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < THREADS / 32) sm.warp_tot[lane] = wi - w;   // exclusive prefix of warp totals
    }
    __syncthreads();
    unsigned int excl = sm.warp_tot[wrp] + incl - v;           // elements before this thread's first bin
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (rank >= excl && rank < excl + h[j]) { sm.sel_bin = PER * tid + j; sm.sel_rank = rank - excl; }
        excl += h[j];
    }
    __syncthreads();
}

// Exact k-th smallest (0-based rank) among the valid elements produced by
// `get(idx, &value) -> bool valid` for idx in [0, n).  NaNs must be excluded by
// `get` (the caller decides what a NaN means).  Requires rank < #valid.
// All threads of the CTA must call; result is returned to all threads.
template <typename Get>
__device__ float select_rank(Smem& sm, int n, unsigned int rank, Get get) {
    const int tid = threadIdx.x;
    uint32_t prefix = 0;       // selected high bits so far
    uint32_t mask = 0;         // which bits of the key are fixed
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = shifts[pass], nb = 1 << widths[pass];
        for (int i = tid; i < kBins; i += kThreads) sm.hist[i] = 0u;
        __syncthreads();
        for (int idx = tid; idx < n; idx += kThreads) {
            float v;
            if (get(idx, &v)) {
                const uint32_t k = float_key(v);
                if ((k & mask) == prefix) atomicAdd(&sm.hist[(k >> sh) & (nb - 1)], 1u);
            }
        }
        __syncthreads();
        pick_bin(sm, rank, nb);
        prefix |= sm.sel_bin << sh;
        mask |= (uint32_t)(nb - 1) << sh;
        rank = sm.sel_rank;
        __syncthreads();
    }
    return key_float(prefix);
}

// Exact k-th smallest (0-based) of n uint32 values < 2^nbits held in shared memory, `vals[i]` = key - base.
// Two (nbits <= 22) or three 11-bit passes from bit nbits - 1 down.  Selecting on the OFFSET inside a narrow
// bracket keeps the first pass well spread: on the raw keys every candidate shares the top bits and all
// shared-memory atomics of the first pass hit one bin (serialised).
__device__ __forceinline__ uint32_t select_rank_offsets(Smem& sm, const unsigned int* vals, int n, unsigned int rank, int nbits) {
    const int tid = threadIdx.x;
    uint32_t prefix = 0, mask = 0;
    int hi = nbits;                                   // bits [hi, 32) are resolved (zero above nbits)
    while (hi > 0) {
        const int w = hi < 11 ? hi : 11, sh = hi - w, nb = 1 << w;
        for (int i = tid; i < nb; i += kThreads) sm.hist[i] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += kThreads) {
            const uint32_t k = vals[i];
            if ((k & mask) == prefix) atomicAdd(&sm.hist[(k >> sh) & (nb - 1)], 1u);
        }
        __syncthreads();
        pick_bin(sm, rank, nb);
        prefix |= sm.sel_bin << sh;
        mask |= (uint32_t)(nb - 1) << sh;
        rank = sm.sel_rank;
        __syncthreads();
        hi = sh;
    }
    return prefix;
}

}  // namespace t3d_select
