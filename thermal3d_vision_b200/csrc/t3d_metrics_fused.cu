// t3d_metrics_fused.cu -- the depth-metric chain of utils/metrics.py:25-69 as ONE persistent kernel for the common
// case (no caller mask, 16-byte aligned, H*W % 4 == 0, prediction planar or the Z channel of an AoS pointmap, GT
// optionally nearest-resampled): the pointmap is read from DRAM once and never copied.
//
// compute_depth_metrics needs two passes over an image -- the medians of gt[mask] and pred[mask] first
// (scale = median(gt) / median(pred), utils/metrics.py:46-48), the per-pixel terms with the scaled prediction second
// (:51-59).  The multi-kernel chain in t3d_metrics.cu materialises planar copies of z and gt between the passes
// (2.0x the algorithmic DRAM traffic).  Here the two passes of an image are work items of the same grid, a few
// images apart in the work queue, so the second pass finds the image's lines in the 126 MB L2:
//
//   X(b, chunk)  count the valid pixels, those below the sampled median brackets, and collect the ~11 % inside them,
//                bucketed by key range (16 buckets per bracket);
//   M(b, stream) exact np.median among the candidates of the one bucket that holds the middle ranks
//                (range narrowing in shared memory; a full radix select over the source if a bracket missed);
//   S(b, chunk)  the per-pixel terms with scale = med_gt / med_pred, fp64 fixed-order partial sums; the last chunk
//                of an image to finish folds the partials into the image's 7 metrics (utils/metrics.py:61-69).
//
// Queue order: X(wave k), M(wave k-1), S(wave k-2), k = 0, 1, ...; a wave is kWave images.  Items only ever wait for
// items EARLIER in the queue, which running CTAs hold: the waits cannot deadlock, whatever the grid size.
// Arithmetic is that of t3d_metrics.cu (same device functions): delta-counts exact, sums fp64 in a fixed order.
#include "t3d_metrics_internal.cuh"

namespace {

using namespace t3d_metrics;

constexpr int kFThreads = 256;
constexpr int kBuckets = 16;             // key-range buckets per bracket
constexpr int kStage = 2048;             // candidates staged per stream per CTA before a flush
constexpr int kSmall = 256;              // exact selection by rank counting at or below this many keys

// counters per image (ints)
enum { C_NV = 0, C_PNAN = 1, C_FALLBACK = 3, C_LT_G = 4, C_LT_P = 5, C_XDONE = 8, C_MDONE = 9, C_SDONE = 10, C_STRIDE = 16 };

struct FusedArgs {
    const float* pred; const float* gt;
    int B, n, H, W, gt_h, gt_w, chunks, wave, median_scaling;
    unsigned int* queue;                 // [1] zero on entry
    int* counters;                       // [B][C_STRIDE] zero on entry
    int* bcount;                         // [B][2][kBuckets] zero on entry
    const unsigned int* bracket;         // [B][2][2] lo, hi (inclusive); lo > hi: none
    unsigned int* cand;                  // [B][2][kBuckets][kBucketCap]
    float* medians;                      // [B][2]
    double* partials;                    // [B][chunks][kNPart]
    float* out; double* out_f64;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_for(const int* p, int target) {      // thread 0 polls; the block follows
    if (threadIdx.x == 0) {
        while (ld_acquire(p) < target) __nanosleep(64);
    }
    __syncthreads();
}
__device__ __forceinline__ void signal(int* p) {                            // call by all threads after the item's writes
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(p, 1);
}

// bucket of a candidate at offset `off` (0 .. w) inside a bracket of width w + 1 keys: monotone in off
__device__ __forceinline__ int bucket_of(unsigned int off, unsigned int w, unsigned int mult) {
    return (w < (unsigned)kBuckets) ? (int)off : (int)__umulhi(off, mult);
}
__device__ __forceinline__ unsigned int bucket_mult(unsigned int w) {       // floor(2^32 * kBuckets / (w + 1)), w >= kBuckets
    return (unsigned int)((((unsigned long long)kBuckets) << 32) / ((unsigned long long)w + 1ull));
}

// one quad (4 consecutive pixels) of the image: GT (optionally nearest-resampled) and predicted depth
template <int PSTRIDE, bool RESAMPLE>
__device__ __forceinline__ void load_quad(const float* __restrict__ gimg, const float4* __restrict__ p4, int q, int W,
                                          const int* __restrict__ stab, float gv[4], float pv[4]) {
    float4 g;
    if (RESAMPLE) {                                   // W % 4 == 0: the quad lies in one row
        const int y = (4 * q) / W, x = 4 * q - y * W;
        const float* row = gimg + stab[W + y];
        g = make_float4(__ldg(row + stab[x]), __ldg(row + stab[x + 1]), __ldg(row + stab[x + 2]), __ldg(row + stab[x + 3]));
    } else {
        g = __ldg(reinterpret_cast<const float4*>(gimg) + q);
    }
    float4 z;
    if (PSTRIDE == 3) {
        const float4 a = __ldg(p4 + 3 * q), bq = __ldg(p4 + 3 * q + 1), c = __ldg(p4 + 3 * q + 2);
        z = make_float4(a.z, bq.y, c.x, c.w);
    } else {
        z = __ldg(p4 + q);
    }
    gv[0] = g.x; gv[1] = g.y; gv[2] = g.z; gv[3] = g.w;
    pv[0] = z.x; pv[1] = z.y; pv[2] = z.z; pv[3] = z.w;
}

struct Smem {
    unsigned int stage[2][kStage];       // staged candidate keys (X); bucket keys / collected keys (M)
    int scount[2];
    int bcnt[2][kBuckets], bbase[2][kBuckets];
    int red[8];
    unsigned int hist[2048];             // M: range histogram (256 bins) / fallback radix histogram (2048 bins)
    unsigned int scan_tmp[kFThreads / 32];
    unsigned int sel_bin, sel_rank;
    double dred[kFThreads / 32][kNPart];
    int flag;
};

// ---- block helpers (kFThreads threads)
__device__ __forceinline__ unsigned int block_sum(Smem& sm, unsigned int v) {
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm.scan_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned int t = 0;
#pragma unroll
    for (int w = 0; w < kFThreads / 32; ++w) t += sm.scan_tmp[w];
    return t;
}
__device__ __forceinline__ unsigned int block_min(Smem& sm, unsigned int v) {
    v = __reduce_min_sync(0xffffffffu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm.scan_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned int t = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < kFThreads / 32; ++w) t = min(t, sm.scan_tmp[w]);
    return t;
}
// bin with cum(bin - 1) <= rank < cum(bin) over sm.hist[0 .. nb), nb a multiple of kFThreads or <= kFThreads:
// results in sm.sel_bin / sm.sel_rank (rank inside the bin).  Requires rank < total.
__device__ __forceinline__ void block_pick_bin(Smem& sm, unsigned int rank, int nb) {
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int per = (nb + kFThreads - 1) / kFThreads;
    unsigned int local = 0;
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < nb) local += sm.hist[i]; }
    unsigned int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) sm.scan_tmp[wrp] = incl;
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < wrp; ++w) before += sm.scan_tmp[w];
    unsigned int excl = before + incl - local;
    if (rank >= excl && rank < excl + local) {
        for (int k = 0; k < per; ++k) {
            const int i = tid * per + k;
            const unsigned int h = (i < nb) ? sm.hist[i] : 0u;
            if (rank < excl + h) { sm.sel_bin = (unsigned)i; sm.sel_rank = rank - excl; break; }
            excl += h;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------ X: count + collect one chunk
template <int PSTRIDE, bool RESAMPLE>
__device__ void item_extract(const FusedArgs& a, Smem& sm, const int* stab, int b, int chunk) {
    const int tid = threadIdx.x, lane = tid & 31, n = a.n;
    const float* __restrict__ gimg = a.gt + (size_t)b * a.gt_h * a.gt_w;
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.pred + (size_t)b * n * PSTRIDE);
    const uint4 br = __ldg(reinterpret_cast<const uint4*>(a.bracket) + b);
    const unsigned int lo_g = br.x, w_g = br.y - br.x, lo_p = br.z, w_p = br.w - br.z;       // lo > hi (no bracket): w wraps,
    const bool has_g = br.y >= br.x, has_p = br.w >= br.z;                                    // masked by has_*
    const unsigned int lo2[2] = {lo_g, lo_p}, w2[2] = {w_g, w_p};
    if (tid < 2) sm.scount[tid] = 0;
    if (tid < 5) sm.red[tid] = 0;
    __syncthreads();

    // staged keys -> the image's bucket lists in global memory (CTA-wide; called with uniform control flow)
    auto flush = [&]() {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int k = min(sm.scount[s], kStage);
            if (k == 0) continue;                                              // uniform: scount is shared
            if (tid < kBuckets) sm.bcnt[s][tid] = 0;
            __syncthreads();
            const unsigned int mult = (w2[s] >= (unsigned)kBuckets) ? bucket_mult(w2[s]) : 0u;
            int my_bucket[kStage / kFThreads], my_slot[kStage / kFThreads];
#pragma unroll
            for (int u = 0; u < kStage / kFThreads; ++u) {
                const int i = tid + u * kFThreads;
                my_bucket[u] = -1;
                if (i < k) {
                    my_bucket[u] = bucket_of(sm.stage[s][i] - lo2[s], w2[s], mult);
                    my_slot[u] = atomicAdd(&sm.bcnt[s][my_bucket[u]], 1);
                }
            }
            __syncthreads();
            if (tid < kBuckets) {
                const int c = sm.bcnt[s][tid];
                int base = c ? atomicAdd(&a.bcount[((size_t)b * 2 + s) * kBuckets + tid], c) : 0;
                if (base + c > kBucketCap) { atomicExch(&a.counters[(size_t)b * C_STRIDE + C_FALLBACK], 1); base = -1; }
                sm.bbase[s][tid] = base;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < kStage / kFThreads; ++u) {
                if (my_bucket[u] < 0) continue;
                const int base = sm.bbase[s][my_bucket[u]];
                if (base >= 0)
                    a.cand[(((size_t)b * 2 + s) * kBuckets + my_bucket[u]) * kBucketCap + base + my_slot[u]] = sm.stage[s][tid + u * kFThreads];
            }
            __syncthreads();
            if (tid == 0) sm.scount[s] = 0;
        }
        __syncthreads();
    };

    int nv = 0, pnan = 0, lt_g = 0, lt_p = 0;
    const int nq = n >> 2, per = (nq + a.chunks - 1) / a.chunks;
    const int q_begin = chunk * per, q_end = min(q_begin + per, nq);
    for (int q0 = q_begin; q0 < q_end; q0 += kFThreads) {                     // uniform trip count
        const int q = q0 + tid;
        if (q < q_end) {
            float gv[4], pv[4];
            load_quad<PSTRIDE, RESAMPLE>(gimg, p4, q, a.W, stab, gv, pv);
            unsigned int kgs[4], kps[4], fg = 0, fp = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned int gb = __float_as_uint(gv[u]), pb = __float_as_uint(pv[u]);
                const bool ok = (gb - 1u) < 0x7f7fffffu;                     // gt > 0 & finite (utils/metrics.py:27)
                const bool pn = pv[u] != pv[u];
                const bool okp = ok && !pn;
                kgs[u] = gb | 0x80000000u;                                   // key of a positive float
                kps[u] = key_of_bits(pb);
                nv += ok; pnan += (ok && pn);
                lt_g += (ok && kgs[u] < lo_g); lt_p += (okp && kps[u] < lo_p);
                fg |= (unsigned)(ok && has_g && (kgs[u] - lo_g) <= w_g) << u;
                fp |= (unsigned)(okp && has_p && (kps[u] - lo_p) <= w_p) << u;
            }
            if (fg) {                       // one shared-memory atomic per thread with candidates
                int slot = atomicAdd(&sm.scount[0], __popc(fg));
#pragma unroll
                for (int u = 0; u < 4; ++u) if (fg & (1u << u)) { if (slot < kStage) sm.stage[0][slot] = kgs[u]; ++slot; }
            }
            if (fp) {
                int slot = atomicAdd(&sm.scount[1], __popc(fp));
#pragma unroll
                for (int u = 0; u < 4; ++u) if (fp & (1u << u)) { if (slot < kStage) sm.stage[1][slot] = kps[u]; ++slot; }
            }
        }
        __syncthreads();
        // an iteration adds at most 4 * kFThreads keys per stream: flush while another one could overflow the stage
        if (sm.scount[0] > kStage - 4 * kFThreads || sm.scount[1] > kStage - 4 * kFThreads) flush();
    }
    flush();
    nv = __reduce_add_sync(0xffffffffu, nv); pnan = __reduce_add_sync(0xffffffffu, pnan);
    lt_g = __reduce_add_sync(0xffffffffu, lt_g); lt_p = __reduce_add_sync(0xffffffffu, lt_p);
    if (lane == 0) {                                   // CTA-level first: one global atomic per counter per CTA
        if (nv) atomicAdd(&sm.red[0], nv);
        if (pnan) atomicAdd(&sm.red[1], pnan);
        if (lt_g) atomicAdd(&sm.red[3], lt_g);
        if (lt_p) atomicAdd(&sm.red[4], lt_p);
    }
    __syncthreads();
    int* c = a.counters + (size_t)b * C_STRIDE;
    if (tid < 5 && sm.red[tid]) atomicAdd(&c[tid < 3 ? tid : tid + 1], sm.red[tid]);
    signal(&c[C_XDONE]);
}

// ------------------------------------------------------------------ M: exact median of one stream of one image
// exact rank-r key (and its successor in sorted order, for the even-count median) among keys[0 .. m) in shared
// memory, m <= kSmall: every thread ranks one key by counting.  *k0 / *k1 valid in all threads afterwards.
__device__ void small_select(Smem& sm, const unsigned int* keys, int m, unsigned int r, unsigned int* k0, unsigned int* nxt) {
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) { sm.sel_bin = 0xffffffffu; sm.sel_rank = 0xffffffffu; }
    __syncthreads();
    if (tid < m) {
        const unsigned int k = keys[tid];
        unsigned int lt = 0, le = 0;
        for (int i = 0; i < m; ++i) { const unsigned int o = keys[i]; lt += o < k; le += o <= k; }
        if (lt <= r && r < le) sm.sel_bin = k;                 // every duplicate of the answer writes the same value
    }
    __syncthreads();
    const unsigned int ans = sm.sel_bin;
    unsigned int nx = 0xffffffffu;
    if (tid < m && keys[tid] > ans) nx = keys[tid];
    *nxt = block_min(sm, nx);
    *k0 = ans;
}

template <int PSTRIDE, bool RESAMPLE>
__device__ float fallback_select(const FusedArgs& a, Smem& sm, const int* stab, int b, int stream, unsigned int rank) {
    // full 3-pass radix select (11 + 11 + 10 bits) over the source image: exact, slow, rare
    const int tid = threadIdx.x, nq = a.n >> 2;
    const float* __restrict__ gimg = a.gt + (size_t)b * a.gt_h * a.gt_w;
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.pred + (size_t)b * a.n * PSTRIDE);
    uint32_t prefix = 0, mask = 0;
    const int shifts[3] = {21, 10, 0}, widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = shifts[pass], nb = 1 << widths[pass];
        for (int i = tid; i < 2048; i += kFThreads) sm.hist[i] = 0u;
        __syncthreads();
        for (int q = tid; q < nq; q += kFThreads) {
            float gv[4], pv[4];
            load_quad<PSTRIDE, RESAMPLE>(gimg, p4, q, a.W, stab, gv, pv);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = (__float_as_uint(gv[u]) - 1u) < 0x7f7fffffu;
                const float v = stream == 0 ? gv[u] : pv[u];
                if (ok && v == v) {
                    const uint32_t k = t3d_select::float_key(v);
                    if ((k & mask) == prefix) atomicAdd(&sm.hist[(k >> sh) & (nb - 1)], 1u);
                }
            }
        }
        __syncthreads();
        block_pick_bin(sm, rank, nb);
        prefix |= sm.sel_bin << sh;
        mask |= (uint32_t)(nb - 1) << sh;
        rank = sm.sel_rank;
        __syncthreads();
    }
    return t3d_select::key_float(prefix);
}

template <int PSTRIDE, bool RESAMPLE>
__device__ void item_median(const FusedArgs& a, Smem& sm, const int* stab, int b, int s) {
    const int tid = threadIdx.x;
    int* c = a.counters + (size_t)b * C_STRIDE;
    wait_for(&c[C_XDONE], a.chunks);
    const int nv = __ldcg(&c[C_NV]);
    float med = 0.f;
    if (nv > 0) {
        const unsigned int r0 = (unsigned)(nv - 1) / 2, r1 = (unsigned)nv / 2;
        if (s == 1 && __ldcg(&c[C_PNAN]) > 0) med = __int_as_float(0x7fc00000);   // NaN in the stream -> median NaN
        else {
            const int lt = __ldcg(&c[C_LT_G + s]);
            const unsigned int lo = a.bracket[4 * b + 2 * s], hi = a.bracket[4 * b + 2 * s + 1];
            const int* bc = a.bcount + ((size_t)b * 2 + s) * kBuckets;
            int nc = 0;
            for (int j = 0; j < kBuckets; ++j) nc += __ldcg(&bc[j]);
            const bool bracket_ok = (__ldcg(&c[C_FALLBACK]) == 0) && hi >= lo && ((int)r0 >= lt) && ((int)r1 < lt + nc);
            float x0, x1;
            if (bracket_ok) {
                // bucket holding rank r0
                int j0 = 0, before = lt;
                for (; j0 < kBuckets; ++j0) { const int k = __ldcg(&bc[j0]); if ((int)r0 < before + k) break; before += k; }
                const unsigned int* src = a.cand + (((size_t)b * 2 + s) * kBuckets + j0) * kBucketCap;
                int m = __ldcg(&bc[j0]);
                unsigned int rr = r0 - (unsigned)before;               // rank inside the current key range
                unsigned int klo = lo, khi = hi;                       // current key range (inclusive), all of bucket j0 inside
                unsigned int k0 = 0, nxt_in = 0xffffffffu;
                bool cached = false;
                if (m <= kStage) {                                     // the bucket fits the stage: read it from L2 once
                    for (int i = tid; i < m; i += kFThreads) sm.stage[0][i] = __ldcg(src + i);
                    cached = true;
                    __syncthreads();
                }
                const unsigned int* keys = cached ? sm.stage[0] : src;
                // narrow [klo, khi] by 256-bin histograms until at most kSmall keys remain (or the range is one key)
                {
                    unsigned int mn = 0xffffffffu, mx = 0u;
                    for (int i = tid; i < m; i += kFThreads) { const unsigned int k = cached ? keys[i] : __ldcg(keys + i); mn = min(mn, k); mx = max(mx, k); }
                    klo = block_min(sm, mn);
                    khi = ~block_min(sm, ~mx);
                }
                int cnt = m;
                unsigned int above = 0xffffffffu;                     // smallest key above the current range seen so far
                while (cnt > kSmall && khi > klo) {
                    const unsigned int w = khi - klo;
                    const unsigned int mult = (w >= 256u) ? (unsigned int)((256ull << 32) / ((unsigned long long)w + 1ull)) : 0u;
                    for (int i = tid; i < 256; i += kFThreads) sm.hist[i] = 0u;
                    __syncthreads();
                    for (int i = tid; i < m; i += kFThreads) {
                        const unsigned int k = cached ? keys[i] : __ldcg(keys + i);
                        if (k >= klo && k <= khi) atomicAdd(&sm.hist[(w < 256u) ? (k - klo) : __umulhi(k - klo, mult)], 1u);
                    }
                    __syncthreads();
                    block_pick_bin(sm, rr, 256);
                    const unsigned int bin = sm.sel_bin;
                    rr = sm.sel_rank;
                    cnt = (int)sm.hist[bin];
                    __syncthreads();
                    // new range: the keys of `bin` (min / max of the members) -- and the smallest key above it
                    unsigned int mn = 0xffffffffu, mx = 0u, ab = 0xffffffffu;
                    for (int i = tid; i < m; i += kFThreads) {
                        const unsigned int k = cached ? keys[i] : __ldcg(keys + i);
                        if (k < klo || k > khi) continue;
                        const unsigned int bb = (w < 256u) ? (k - klo) : __umulhi(k - klo, mult);
                        if (bb == bin) { mn = min(mn, k); mx = max(mx, k); }
                        else if (bb > bin) ab = min(ab, k);
                    }
                    const unsigned int nlo = block_min(sm, mn), nhi = ~block_min(sm, ~mx);
                    above = min(above, block_min(sm, ab));
                    klo = nlo; khi = nhi;
                }
                if (khi == klo) {                                      // one distinct key left (cnt copies of it)
                    k0 = klo;
                    nxt_in = ((unsigned)cnt > rr + 1u) ? klo : above;  // a duplicate past rank rr, else the next larger key
                } else {
                    // collect the <= kSmall keys of the range and rank them
                    if (tid == 0) sm.flag = 0;
                    __syncthreads();
                    for (int i = tid; i < m; i += kFThreads) {
                        const unsigned int k = cached ? keys[i] : __ldcg(keys + i);
                        if (k >= klo && k <= khi) { const int slot = atomicAdd(&sm.flag, 1); if (slot < kSmall) sm.stage[1][slot] = k; }
                    }
                    __syncthreads();
                    const int mm = min(sm.flag, kSmall);
                    unsigned int nx;
                    small_select(sm, sm.stage[1], mm, rr, &k0, &nx);
                    // successor of the rank-rr element: a duplicate / the next key inside the range, else the smallest above it
                    unsigned int le = 0;
                    if (tid < mm) le = sm.stage[1][tid] <= k0;
                    le = block_sum(sm, le);
                    nxt_in = (le > rr + 1u) ? k0 : min(nx, above);
                }
                x0 = t3d_select::key_float(k0);
                x1 = x0;
                if (r1 != r0 && nxt_in != k0) {
                    unsigned int nk = nxt_in;
                    if (nk == 0xffffffffu) {                           // the successor lives in a later bucket: its minimum
                        for (int j = j0 + 1; j < kBuckets && nk == 0xffffffffu; ++j) {
                            const int mj = __ldcg(&bc[j]);
                            if (mj == 0) continue;
                            const unsigned int* sj = a.cand + (((size_t)b * 2 + s) * kBuckets + j) * kBucketCap;
                            unsigned int mn = 0xffffffffu;
                            for (int i = tid; i < mj; i += kFThreads) mn = min(mn, __ldcg(sj + i));
                            nk = block_min(sm, mn);
                        }
                    }
                    x1 = t3d_select::key_float(nk);
                }
            } else {                                              // fallback: full radix select, same result
                x0 = fallback_select<PSTRIDE, RESAMPLE>(a, sm, stab, b, s, r0);
                x1 = (r1 == r0) ? x0 : fallback_select<PSTRIDE, RESAMPLE>(a, sm, stab, b, s, r1);
            }
            med = (r1 == r0) ? x0 : __fmul_rn(__fadd_rn(x0, x1), 0.5f);   // np.median: fp32 mean of the middles
        }
    }
    if (tid == 0) a.medians[2 * b + s] = med;
    signal(&c[C_MDONE]);
}

// ------------------------------------------------------------------ S: per-pixel terms of one chunk (+ finalize)
template <int PSTRIDE, bool RESAMPLE>
__device__ void item_sums(const FusedArgs& a, Smem& sm, const int* stab, int b, int chunk) {
    const int tid = threadIdx.x, n = a.n;
    int* c = a.counters + (size_t)b * C_STRIDE;
    float s = 1.0f;
    if (a.median_scaling) {
        wait_for(&c[C_MDONE], 2);
        if (__ldcg(&c[C_NV]) > 0) s = __fdiv_rn(__ldcg(&a.medians[2 * b]), __ldcg(&a.medians[2 * b + 1]));   // utils/metrics.py:47
    } else {
        wait_for(&c[C_XDONE], a.chunks);
    }
    const float* __restrict__ gimg = a.gt + (size_t)b * a.gt_h * a.gt_w;
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.pred + (size_t)b * n * PSTRIDE);
    double acc[4] = {0, 0, 0, 0};
    int cnt[3] = {0, 0, 0};
    const int nq = n >> 2, per = (nq + a.chunks - 1) / a.chunks;
    const int q0 = chunk * per, q1 = min(q0 + per, nq);
    const float qnan = __int_as_float(0x7fc00000);
    for (int q = q0 + tid; q < q1; q += 2 * kFThreads) {
        const int q2 = q + kFThreads;
        float gt[8], za8[8];
        load_quad<PSTRIDE, RESAMPLE>(gimg, p4, q, a.W, stab, gt, za8);
        if (q2 < q1) load_quad<PSTRIDE, RESAMPLE>(gimg, p4, q2, a.W, stab, gt + 4, za8 + 4);
        else {
#pragma unroll
            for (int e = 4; e < 8; ++e) { gt[e] = qnan; za8[e] = qnan; }
        }
        float pr[8];
        float accf[4] = {0.f, 0.f, 0.f, 0.f};
        float lowest = 1.0f;                                                         // NaN-propagating min of the scaled predictions
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const bool ok = (__float_as_uint(gt[e]) - 1u) < 0x7f7fffffu;             // gt > 0 & finite (:27); else the "invalid" marker
            gt[e] = ok ? gt[e] : qnan;
            pr[e] = __fmul_rn(za8[e], s);                                            // pred *= scale   (:48)
            const float m = ok ? pr[e] : 1.0f;                                       // unselected pixels do not matter
            asm("min.NaN.f32 %0, %0, %1;" : "+f"(lowest) : "f"(m));
        }
        if (lowest > 0.f) {                                                          // all selected predictions positive (not NaN)
#pragma unroll
            for (int e = 0; e < 8; ++e) metric_terms_fast(gt[e], pr[e], accf, cnt);
        } else {
#pragma unroll 1
            for (int e = 0; e < 8; ++e) metric_terms<false>(gt[e], za8[e], s, accf, cnt);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += (double)accf[k];
    }
    double v[kNPart] = {acc[0], acc[1], acc[2], acc[3], (double)cnt[0], (double)cnt[1], (double)cnt[2], 0.0};
#pragma unroll
    for (int k = 0; k < kNPart - 1; ++k) v[k] = warp_sum(v[k]);
    const int lane = tid & 31, wrp = tid >> 5;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kNPart; ++k) sm.dred[wrp][k] = v[k];
    }
    __syncthreads();
    if (tid < kNPart) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < kFThreads / 32; ++w) t += sm.dred[w][tid];
        a.partials[((size_t)b * a.chunks + chunk) * kNPart + tid] = t;
    }
    // the last chunk of the image to finish folds the partials (fixed order) into the metrics (utils/metrics.py:61-69)
    __threadfence();
    __syncthreads();
    if (tid == 0) sm.flag = (atomicAdd(&c[C_SDONE], 1) == a.chunks - 1);
    __syncthreads();
    if (!sm.flag) return;
    __threadfence();
    const int k = tid >> 5;                            // one warp per metric: lane owns chunks lane, lane + 32, ...
    if (k < kNPart) {
        double t = 0;
        for (int ch = lane; ch < a.chunks; ch += 32) t += __ldcg(&a.partials[((size_t)b * a.chunks + ch) * kNPart + k]);
        t = warp_sum(t);
        if (lane == 0) {
            const int nv = __ldcg(&c[C_NV]);
            const double dq = __longlong_as_double(0x7ff8000000000000LL);
            double r;
            if (k == 7) r = (double)nv;
            else if (nv == 0) r = (k < 4) ? dq : 0.0;                                    // utils/metrics.py:34-43
            else if (k < 2) r = (double)(float)(t / nv);
            else if (k < 4) r = (double)sqrtf((float)(t / nv));                          // np.sqrt(np.mean(.)) in fp32
            else r = t / nv;                                                             // (thresh < t).mean() -> fp64
            a.out[(size_t)b * 8 + k] = (float)r;
            if (a.out_f64) a.out_f64[(size_t)b * 8 + k] = r;
        }
    }
}

// ------------------------------------------------------------------ the persistent kernel
template <int PSTRIDE, bool RESAMPLE>
__global__ void __launch_bounds__(kFThreads, 3) metrics_fused_kernel(const FusedArgs a) {
    __shared__ Smem sm;
    __shared__ int stab[RESAMPLE ? kResampleMaxDim : 1];          // cv2 INTER_NEAREST source column (W) / row offset (H)
    __shared__ int s_task;
    const int tid = threadIdx.x;
    if (RESAMPLE) {     // utils/evaluate_depth_metrics.py:321-323: sx = min(floor(x * gw / W), gw - 1), same for rows
        const double fx = (double)a.gt_w / (double)a.W, fy = (double)a.gt_h / (double)a.H;
        for (int i = tid; i < a.W + a.H; i += kFThreads) {
            if (i < a.W) stab[i] = min((int)floor(__dmul_rn((double)i, fx)), a.gt_w - 1);
            else stab[i] = min((int)floor(__dmul_rn((double)(i - a.W), fy)), a.gt_h - 1) * a.gt_w;
        }
    }
    const int nwaves = (a.B + a.wave - 1) / a.wave;
    const int m_items = a.median_scaling ? 2 : 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_task = (int)atomicAdd(a.queue, 1u);
        __syncthreads();
        int t = s_task;
        // decode: slot k holds X(wave k), M(wave k - 1), S(wave k - 2)
        int kind = -1, b = 0, sub = 0;
        for (int k = 0; k < nwaves + 2 && kind < 0; ++k) {
            for (int ph = 0; ph < 3 && kind < 0; ++ph) {
                const int w = k - ph;
                if (w < 0 || w >= nwaves) continue;
                const int nimg = min(a.wave, a.B - w * a.wave);
                const int per = (ph == 1) ? m_items : a.chunks;
                const int cnt = nimg * per;
                if (t < cnt) { kind = ph; b = w * a.wave + t / per; sub = t % per; }
                else t -= cnt;
            }
        }
        if (kind < 0) break;
        if (kind == 0) item_extract<PSTRIDE, RESAMPLE>(a, sm, stab, b, sub);
        else if (kind == 1) item_median<PSTRIDE, RESAMPLE>(a, sm, stab, b, sub);
        else item_sums<PSTRIDE, RESAMPLE>(a, sm, stab, b, sub);
    }
}

}  // namespace

namespace t3d_metrics {

size_t fused_ws_bytes(int B, int chunks) {
    size_t off = 0;
    off += t3d_align_up(256, 256);                                                   // queue
    off += t3d_align_up((size_t)B * C_STRIDE * sizeof(int), 256);                    // counters
    off += t3d_align_up((size_t)B * 2 * kBuckets * sizeof(int), 256);                // bucket counts
    off += t3d_align_up((size_t)B * 2 * kBuckets * kBucketCap * sizeof(unsigned int), 256);
    return off;
}

int launch_fused(const float* pred, int pred_stride, const float* gt, int gt_h, int gt_w, int B, int H, int W,
                 int median_scaling, const unsigned int* bracket, float* medians, double* partials, int chunks,
                 float* out, double* out_f64, void* ws, cudaStream_t st) {
    FusedArgs a;
    char* p = reinterpret_cast<char*>(ws);
    size_t off = 0;
    a.queue = reinterpret_cast<unsigned int*>(p + off); off += 256;
    a.counters = reinterpret_cast<int*>(p + off); off += t3d_align_up((size_t)B * C_STRIDE * sizeof(int), 256);
    a.bcount = reinterpret_cast<int*>(p + off); off += t3d_align_up((size_t)B * 2 * kBuckets * sizeof(int), 256);
    const size_t zero_bytes = off;
    a.cand = reinterpret_cast<unsigned int*>(p + off);
    T3D_CUDA(cudaMemsetAsync(p, 0, zero_bytes, st));
    a.pred = pred; a.gt = gt; a.B = B; a.n = H * W; a.H = H; a.W = W; a.gt_h = gt_h; a.gt_w = gt_w;
    a.chunks = chunks; a.median_scaling = median_scaling ? 1 : 0;
    static const int wave = [] { const char* e = getenv("T3D_METRIC_WAVE"); const int v = e ? atoi(e) : 8; return v < 1 ? 1 : v; }();
    a.wave = wave;
    a.bracket = bracket; a.medians = medians; a.partials = partials; a.out = out; a.out_f64 = out_f64;
    const bool resample = (gt_h != H) || (gt_w != W);
    const long long items = (long long)B * (2 * chunks + 2);
    static const int ctas_per_sm = [] { const char* e = getenv("T3D_METRIC_CTAS"); const int v = e ? atoi(e) : 3; return v < 1 ? 1 : v; }();
    int grid = t3d_sm_count() * ctas_per_sm;
    if ((long long)grid > items) grid = (int)items;
#define T3D_FUSED(PS_, RS_) T3D_LAUNCH("metrics_fused_kernel", st, (metrics_fused_kernel<PS_, RS_><<<grid, kFThreads, 0, st>>>(a)))
    if (pred_stride == 3) { if (resample) T3D_FUSED(3, true); else T3D_FUSED(3, false); }
    else { if (resample) T3D_FUSED(1, true); else T3D_FUSED(1, false); }
#undef T3D_FUSED
    return T3D_OK;
}

}  // namespace t3d_metrics
