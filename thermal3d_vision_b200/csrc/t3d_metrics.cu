// t3d_metrics.cu -- pointmap -> depth and the depth-metric reductions (sm_100a).
//
// Replaces /root/reference/utils/metrics.py:4-69 (compute_depth_metrics, 7
// metrics), utils/evaluate_depth_metrics.py:20-80 (3-metric variant, a subset),
// the z-extraction `pointmap[..., 2]` (utils/metrics.py:121,
// thermal_dustr_inference.py:133-134, scripts/pseudo_gt.py:115-116), the GT
// nearest resample (utils/evaluate_depth_metrics.py:320-323) and the median
// focal estimate of scripts/pseudo_gt.py:151-184.
//
// Per image: mask = gt > 0 & finite (or the caller's mask); exact float32
// medians of gt[mask] and pred[mask] by radix select; pred *= med_gt/med_pred;
// then every per-pixel term in float32 exactly as numpy evaluates it
// (IEEE div/mul, no FMA), summed in fp64 in a fixed order (deterministic).
#include "t3d_common.cuh"
#include "t3d_select.cuh"

namespace {

constexpr int kChunkThreads = 256;
constexpr int kNPart = 8;   // abs_rel, sq_rel, sq, log2, a1, a2, a3, (pad)

// Batched exact medians: multi-CTA 3-pass radix select (11 + 11 + 10 bits of the monotone key).
// Two streams per image (0 = gt, 1 = pred) and two targets per stream (A = rank (n-1)/2, B = rank n/2;
// np.median averages them).  Histogram kernels run on (chunks x B) CTAs with shared-memory histograms
// and warp-aggregated atomics; tiny pick kernels scan the merged histograms between passes.
constexpr int kSelBins = 2048;

struct SelState {           // per (image, stream)
    unsigned int prefix[2]; // selected high bits of target A / B
    unsigned int rank[2];   // remaining rank inside the prefix
};

struct MetricsWs {
    float* vz; float* vg;
    int* counters;          // [B][4]: n_valid, pred_nan, gt_nan, pad
    unsigned int* hist;     // [B][2 streams][2 targets][kSelBins]
    SelState* state;        // [B][2]
    float* scale;           // [B]
    double* partials;       // [B][chunks][kNPart]
    size_t zero_bytes;      // counters + hist are contiguous: one memset
    size_t total;
};

MetricsWs metrics_ws(void* base, int B, int n, int chunks) {
    MetricsWs w;
    size_t off = 0;
    char* p = reinterpret_cast<char*>(base);
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += t3d_align_up(bytes, 256); return r; };
    w.counters = reinterpret_cast<int*>(take((size_t)B * 4 * sizeof(int)));
    w.hist = reinterpret_cast<unsigned int*>(take((size_t)B * 4 * kSelBins * sizeof(unsigned int)));
    w.zero_bytes = off;
    w.vz = reinterpret_cast<float*>(take((size_t)B * n * 4));
    w.vg = reinterpret_cast<float*>(take((size_t)B * n * 4));
    w.state = reinterpret_cast<SelState*>(take((size_t)B * 2 * sizeof(SelState)));
    w.scale = reinterpret_cast<float*>(take((size_t)B * sizeof(float)));
    w.partials = reinterpret_cast<double*>(take((size_t)B * chunks * kNPart * sizeof(double)));
    w.total = off;
    return w;
}

int chunks_for(int n) { return max(1, min(96, (n + 4095) / 4096)); }

// one shared-memory histogram increment, aggregated over the lanes of the warp that hit the same bin
// (depth values crowd into a few exponent bins: without aggregation same-address atomics serialise)
__device__ __forceinline__ void hist_add(unsigned int* h, bool on, unsigned int bin) {
    const unsigned int key = on ? bin : 0xffffffffu;
    const unsigned int peers = __match_any_sync(0xffffffffu, key);
    if (on && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[bin], (unsigned)__popc(peers));
}

__device__ __forceinline__ void merge_hist(const unsigned int* sh, unsigned int* gh, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const unsigned int v = sh[i];
        if (v) atomicAdd(&gh[i], v);
    }
}

// ------------------------------------------------------------------ M1: extract (K5) + first select pass
// pred element (b, i) lives at pred[(b*n + i) * pred_stride + pred_offset]: stride 3 / offset 2 reads
// the Z channel of an AoS pointmap in place (depth is never materialised by the caller).
// Invalid pixels are stored as vg = NaN (a *selected* NaN GT makes every metric NaN anyway: gt_nan counter).
__global__ void __launch_bounds__(kChunkThreads)
depth_extract_kernel(const float* __restrict__ pred, int pred_stride, int pred_offset,
                     const float* __restrict__ gt, int gt_h, int gt_w, const unsigned char* __restrict__ mask,
                     int H, int W, float* __restrict__ vz, float* __restrict__ vg,
                     int* __restrict__ counters, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[2][kSelBins];
    const int b = blockIdx.y, n = H * W;
    for (int i = threadIdx.x; i < 2 * kSelBins; i += kChunkThreads) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const bool resample = (gt_h != H) || (gt_w != W);
    const double fx = (double)gt_w / (double)W, fy = (double)gt_h / (double)H;
    const float* g = gt + (size_t)b * gt_h * gt_w;
    const float* p = pred + (size_t)b * n * pred_stride + pred_offset;
    int nv = 0, pnan = 0, gnan = 0;
    const int span = gridDim.x * kChunkThreads;
    for (int i0 = blockIdx.x * kChunkThreads; i0 < n; i0 += span) {      // warp-uniform trip count
        const int i = i0 + threadIdx.x;
        bool ok = false;
        float gv = 0.f, pv = 0.f;
        if (i < n) {
            if (resample) {   // cv2 INTER_NEAREST (utils/evaluate_depth_metrics.py:321-323)
                const int y = i / W, x = i - y * W;
                const int sx = min((int)floor(__dmul_rn((double)x, fx)), gt_w - 1);
                const int sy = min((int)floor(__dmul_rn((double)y, fy)), gt_h - 1);
                gv = __ldg(g + (size_t)sy * gt_w + sx);
            } else {
                gv = __ldg(g + i);
            }
            pv = __ldg(p + (size_t)i * pred_stride);
            ok = mask ? (mask[(size_t)b * n + i] != 0) : (gv > 0.f && isfinite(gv));   // utils/metrics.py:27
            vz[(size_t)b * n + i] = pv;
            vg[(size_t)b * n + i] = ok ? gv : __int_as_float(0x7fc00000);
            if (ok) { ++nv; pnan += isnan(pv); gnan += isnan(gv); }
        }
        hist_add(sh[0], ok && !isnan(gv), t3d_select::float_key(gv) >> 21);
        hist_add(sh[1], ok && !isnan(pv), t3d_select::float_key(pv) >> 21);
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    pnan = __reduce_add_sync(0xffffffffu, pnan);
    gnan = __reduce_add_sync(0xffffffffu, gnan);
    if ((threadIdx.x & 31) == 0) {
        if (nv) atomicAdd(&counters[4 * b], nv);
        if (pnan) atomicAdd(&counters[4 * b + 1], pnan);
        if (gnan) atomicAdd(&counters[4 * b + 2], gnan);
    }
    __syncthreads();
    unsigned int* gh = hist + (size_t)b * 4 * kSelBins;
    merge_hist(sh[0], gh, kSelBins);                      // stream 0, target A
    merge_hist(sh[1], gh + 2 * kSelBins, kSelBins);       // stream 1, target A
}

// ------------------------------------------------------------------ select: pick kernel (one CTA per image)
// PASS 0/1/2 = after the histogram of bits [31:21] / [20:10] / [9:0].
__device__ __forceinline__ void scan_pick(const unsigned int* __restrict__ gh, unsigned int* sbuf, unsigned int* wtot,
                                          unsigned int rank, int nb, unsigned int* out_bin, unsigned int* out_rank) {
    // 256 threads x 8 bins; deterministic prefix scan
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    unsigned int loc[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { loc[k] = (8 * tid + k < nb) ? gh[8 * tid + k] : 0u; sum += loc[k]; }
    unsigned int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wtot[wrp] = incl;
    __syncthreads();
    if (tid == 0) { unsigned int a = 0; for (int w = 0; w < 8; ++w) { const unsigned int t = wtot[w]; wtot[w] = a; a += t; } }
    __syncthreads();
    unsigned int c = wtot[wrp] + incl - sum;
    if (rank >= c && rank < c + sum) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (rank < c + loc[k]) { sbuf[0] = 8 * tid + k; sbuf[1] = rank - c; break; }
            c += loc[k];
        }
    }
    __syncthreads();
    *out_bin = sbuf[0]; *out_rank = sbuf[1];
    __syncthreads();
}

template <int PASS>
__global__ void __launch_bounds__(256) select_pick_kernel(unsigned int* __restrict__ hist, SelState* __restrict__ state,
                                                          const int* __restrict__ counters, int median_scaling,
                                                          float* __restrict__ scale, float* __restrict__ out_medians) {
    __shared__ unsigned int sbuf[2], wtot[8];
    __shared__ float med[2];
    const int b = blockIdx.x;
    const int nv = counters[4 * b];
    constexpr int SH = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
    constexpr int NB = (PASS == 2) ? 1024 : 2048;
    unsigned int* gh = hist + (size_t)b * 4 * kSelBins;
    if (nv > 0 && median_scaling) {
        for (int s = 0; s < 2; ++s) {
            SelState st = (PASS == 0) ? SelState{{0u, 0u}, {(unsigned)(nv - 1) / 2, (unsigned)nv / 2}} : state[2 * b + s];
            if (counters[4 * b + (s == 0 ? 2 : 1)] > 0) continue;        // NaN in the stream -> median NaN
            const bool shared_hist = (PASS == 0) || (st.prefix[0] == st.prefix[1]);
            unsigned int bin, rk;
            scan_pick(gh + (2 * s) * kSelBins, sbuf, wtot, st.rank[0], NB, &bin, &rk);
            st.prefix[0] |= bin << SH; st.rank[0] = rk;
            scan_pick(gh + (2 * s + (shared_hist ? 0 : 1)) * kSelBins, sbuf, wtot, st.rank[1], NB, &bin, &rk);
            st.prefix[1] |= bin << SH; st.rank[1] = rk;
            if (threadIdx.x == 0) {
                state[2 * b + s] = st;
                if (PASS == 2) {                                          // np.median: fp32 mean of the two middles
                    const float a = t3d_select::key_float(st.prefix[0]), c = t3d_select::key_float(st.prefix[1]);
                    med[s] = (st.prefix[0] == st.prefix[1]) ? a : __fmul_rn(__fadd_rn(a, c), 0.5f);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * kSelBins; i += 256) gh[i] = 0u;     // ready for the next pass / next call
    if (PASS == 2 && threadIdx.x == 0) {
        float s = 1.0f, mg = 0.f, mp = 0.f;
        if (nv > 0 && median_scaling) {
            const float qnan = __int_as_float(0x7fc00000);
            mg = counters[4 * b + 2] > 0 ? qnan : med[0];
            mp = counters[4 * b + 1] > 0 ? qnan : med[1];
            s = __fdiv_rn(mg, mp);                                        // utils/metrics.py:47
        }
        scale[b] = s;
        if (out_medians) { out_medians[2 * b] = mg; out_medians[2 * b + 1] = mp; }
    }
}

// ------------------------------------------------------------------ select: histogram kernel for passes 1 and 2
template <int PASS>
__global__ void __launch_bounds__(kChunkThreads)
select_hist_kernel(const float* __restrict__ vz, const float* __restrict__ vg, const SelState* __restrict__ state,
                   const int* __restrict__ counters, int n, int median_scaling, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[4][kSelBins];
    const int b = blockIdx.y;
    if (!median_scaling || counters[4 * b] == 0) return;
    constexpr int SH = (PASS == 1) ? 10 : 0;
    constexpr unsigned int DM = (PASS == 1) ? 2047u : 1023u;
    constexpr unsigned int FIXED = (PASS == 1) ? 0xffe00000u : 0xfffffc00u;
    for (int i = threadIdx.x; i < 4 * kSelBins; i += kChunkThreads) (&sh[0][0])[i] = 0u;
    const SelState sg = state[2 * b], sp = state[2 * b + 1];
    const bool two_g = sg.prefix[0] != sg.prefix[1], two_p = sp.prefix[0] != sp.prefix[1];
    __syncthreads();
    const float* g = vg + (size_t)b * n;
    const float* z = vz + (size_t)b * n;
    const int span = gridDim.x * kChunkThreads;
    for (int i0 = blockIdx.x * kChunkThreads; i0 < n; i0 += span) {
        const int i = i0 + threadIdx.x;
        float gv = __int_as_float(0x7fc00000), pv = 0.f;
        if (i < n) { gv = __ldg(g + i); pv = __ldg(z + i); }
        const bool ok = !isnan(gv);
        const unsigned int kg = t3d_select::float_key(gv), kp = t3d_select::float_key(pv);
        hist_add(sh[0], ok && (kg & FIXED) == sg.prefix[0], (kg >> SH) & DM);
        if (two_g) hist_add(sh[1], ok && (kg & FIXED) == sg.prefix[1], (kg >> SH) & DM);
        const bool okp = ok && !isnan(pv);
        hist_add(sh[2], okp && (kp & FIXED) == sp.prefix[0], (kp >> SH) & DM);
        if (two_p) hist_add(sh[3], okp && (kp & FIXED) == sp.prefix[1], (kp >> SH) & DM);
    }
    __syncthreads();
    merge_hist(&sh[0][0], hist + (size_t)b * 4 * kSelBins, 4 * kSelBins);
}

// ------------------------------------------------------------------ M3: per-pixel terms
__device__ __forceinline__ float np_maximum(float a, float b) { return (isnan(a) || isnan(b)) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

__global__ void __launch_bounds__(kChunkThreads)
metrics_sum_kernel(const float* __restrict__ vz, const float* __restrict__ vg,
                   const float* __restrict__ scale, int n, int chunks, double* __restrict__ partials) {
    __shared__ double red[kChunkThreads / 32][kNPart];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const float s = scale[b];
    const int per = (n + chunks - 1) / chunks;
    const int i0 = chunk * per, i1 = min(i0 + per, n);
    float accf[4] = {0.f, 0.f, 0.f, 0.f};        // short fp32 runs (<= 16 terms) folded into fp64
    double acc[4] = {0, 0, 0, 0};
    int cnt[3] = {0, 0, 0};
    int run = 0;
    for (int i = i0 + threadIdx.x; i < i1; i += kChunkThreads) {
        const float gt = __ldg(vg + (size_t)b * n + i);
        if (isnan(gt)) continue;                                                 // invalid pixel marker
        const float pr = __fmul_rn(__ldg(vz + (size_t)b * n + i), s);            // pred *= scale   (:48)
        const float q = __fdiv_rn(gt, pr);
        const float th = np_maximum(q, __fdiv_rn(pr, gt));                       // :51 (exact IEEE: counts are exact)
        cnt[0] += th < 1.25f; cnt[1] += th < 1.5625f; cnt[2] += th < 1.953125f;  // :52-54
        const float d = __fsub_rn(gt, pr);
        const float d2 = __fmul_rn(d, d);
        const float rg = __frcp_rn(gt);
        accf[0] += fabsf(d) * rg;                                                // :56  |gt - pred| / gt
        accf[1] += d2 * rg;                                                      // :57
        accf[2] += d2;                                                           // :58
        const float dl = __fsub_rn(logf(gt), logf(pr));
        accf[3] += dl * dl;                                                      // :59
        if (++run == 16) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { acc[k] += (double)accf[k]; accf[k] = 0.f; }
            run = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += (double)accf[k];
    double v[kNPart] = {acc[0], acc[1], acc[2], acc[3], (double)cnt[0], (double)cnt[1], (double)cnt[2], 0.0};
#pragma unroll
    for (int k = 0; k < kNPart - 1; ++k) v[k] = warp_sum(v[k]);
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kNPart; ++k) red[wrp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < kNPart) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < kChunkThreads / 32; ++w) t += red[w][threadIdx.x];
        partials[((size_t)b * chunks + chunk) * kNPart + threadIdx.x] = t;
    }
}

// ------------------------------------------------------------------ M4: finalize
// out[b] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, n_valid  (float32 like numpy's results;
// out_f64 keeps a_k = count / n in fp64 as numpy returns them)
__global__ void metrics_finalize_kernel(const double* __restrict__ partials, const int* __restrict__ counters,
                                        int chunks, float* __restrict__ out, double* __restrict__ out_f64) {
    const int b = blockIdx.x, k = threadIdx.x;
    if (k >= kNPart) return;
    double s = 0;
    for (int c = 0; c < chunks; ++c) s += partials[((size_t)b * chunks + c) * kNPart + k];
    const int nv = counters[4 * b];
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (counters[4 * b + 2] > 0 && k < 4) s = qnan;                              // a selected NaN GT poisons every mean
    double r;
    if (k == 7) r = (double)nv;
    else if (nv == 0) r = (k < 4) ? qnan : 0.0;                                  // utils/metrics.py:34-43
    else if (k < 2) r = (double)(float)(s / nv);
    else if (k < 4) r = (double)sqrtf((float)(s / nv));                          // np.sqrt(np.mean(.)) in fp32
    else r = s / nv;                                                             // (thresh < t).mean() -> fp64
    out[(size_t)b * 8 + k] = (float)r;
    if (out_f64) out_f64[(size_t)b * 8 + k] = r;
}

// ------------------------------------------------------------------ intrinsics: median focal estimate
// scripts/pseudo_gt.py:151-184: fx = median((u - W/2) / (X/Z)), fy = median((v - H/2) / (Y/Z)) over Z > 0, fp32
__global__ void __launch_bounds__(t3d_select::kThreads, 1)
focal_estimate_kernel(const float* __restrict__ pointmap, const float* __restrict__ depth, int H, int W,
                      double* __restrict__ out_K) {
    __shared__ t3d_select::Smem sm;
    __shared__ int s_cnt[3];
    const int b = blockIdx.x, n = H * W;
    const float* pm = pointmap + (size_t)b * n * 3;
    const float* dz = depth ? depth + (size_t)b * n : nullptr;
    auto Z = [&](int i) { return dz ? dz[i] : pm[(size_t)i * 3 + 2]; };
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const float hw = (float)((double)W / 2.0), hh = (float)((double)H / 2.0);
    auto ratio = [&](int i, int axis) {
        const float z = Z(i);
        const float c = pm[(size_t)i * 3 + axis];
        const int y = i / W, x = i - y * W;
        // numpy: (u - W/2) is float64 (int64 - python float), X/Z float32 -> division in float64
        const double num = axis == 0 ? ((double)x - (double)W / 2.0) : ((double)y - (double)H / 2.0);
        return num / (double)__fdiv_rn(c, z);
    };
    (void)hw; (void)hh;
    // float64 medians: select on the fp32-rounded keys is not exact for fp64 data, so do a
    // 2-level refinement: this helper is a convenience, not a hot path -> simple O(passes * n) search
    // on the monotone 64-bit key, 16 bits per pass.
    double med[2];
    for (int axis = 0; axis < 2; ++axis) {
        int nv = 0, nn = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (Z(i) > 0.f) { ++nv; nn += isnan(ratio(i, axis)); }
        }
        atomicAdd(&s_cnt[0], nv); atomicAdd(&s_cnt[1], nn);
        __syncthreads();
        const int total = s_cnt[0], nans = s_cnt[1];
        __syncthreads();
        if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
        __syncthreads();
        if (total == 0 || nans > 0) { med[axis] = __longlong_as_double(0x7ff8000000000000LL); continue; }
        auto key64 = [&](double d) {
            unsigned long long u = (unsigned long long)__double_as_longlong(d);
            return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
        };
        double vals[2];
        const unsigned int ranks[2] = {(unsigned)(total - 1) / 2, (unsigned)total / 2};
        for (int rr = 0; rr < 2; ++rr) {
            if (rr == 1 && ranks[1] == ranks[0]) { vals[1] = vals[0]; break; }
            unsigned long long prefix = 0, fixed = 0;
            unsigned int rank = ranks[rr];
            for (int pass = 0; pass < 6; ++pass) {      // 11,11,11,11,11,9 bits
                const int width = pass < 5 ? 11 : 9;
                const int sh = 64 - 11 * pass - width;
                for (int i = threadIdx.x; i < t3d_select::kBins; i += blockDim.x) sm.hist[i] = 0u;
                __syncthreads();
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    if (Z(i) > 0.f) {
                        const unsigned long long k = key64(ratio(i, axis));
                        if ((k & fixed) == prefix) atomicAdd(&sm.hist[(k >> sh) & ((1u << width) - 1)], 1u);
                    }
                }
                __syncthreads();
                t3d_select::pick_bin(sm, rank, 1 << width);
                prefix |= (unsigned long long)sm.sel_bin << sh;
                fixed |= (unsigned long long)((1u << width) - 1) << sh;
                rank = sm.sel_rank;
                __syncthreads();
            }
            const unsigned long long u = (prefix >> 63) ? (prefix & 0x7fffffffffffffffULL) : ~prefix;
            vals[rr] = __longlong_as_double((long long)u);
        }
        med[axis] = (ranks[0] == ranks[1]) ? vals[0] : (vals[0] + vals[1]) / 2.0;   // np.median: mean of the two
    }
    if (threadIdx.x == 0) {
        double* K = out_K + (size_t)b * 9;
        K[0] = med[0]; K[1] = 0; K[2] = (double)W / 2.0;
        K[3] = 0; K[4] = med[1]; K[5] = (double)H / 2.0;
        K[6] = 0; K[7] = 0; K[8] = 1;
    }
}

// pointmap -> dense depth map (when a caller really wants the [H,W] array, e.g. np.save in inference)
__global__ void __launch_bounds__(256) pointmap_to_depth_kernel(const float* __restrict__ pm, float* __restrict__ depth, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        depth[i] = __ldg(pm + i * 3 + 2);
}

// EXTENSION (not in the reference): u = fx X/Z + cx, v = fy Y/Z + cy
__global__ void __launch_bounds__(256) project_points_kernel(const float* __restrict__ pm, float fx, float fy, float cx,
                                                             float cy, float* __restrict__ uv, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float X = pm[i * 3], Y = pm[i * 3 + 1], Z = pm[i * 3 + 2];
        uv[2 * i] = __fadd_rn(__fmul_rn(fx, __fdiv_rn(X, Z)), cx);
        uv[2 * i + 1] = __fadd_rn(__fmul_rn(fy, __fdiv_rn(Y, Z)), cy);
    }
}

}  // namespace

extern "C" {

size_t t3d_depth_metrics_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    return metrics_ws(nullptr, B, H * W, chunks_for(H * W)).total;
}

int t3d_depth_metrics(const float* pred, int pred_stride, int pred_offset,
                      const float* gt, int gt_h, int gt_w, const unsigned char* mask,
                      int B, int H, int W, int median_scaling,
                      float* out, double* out_f64, float* out_medians,
                      void* workspace, size_t workspace_bytes, void* stream) {
    T3D_REQUIRE(pred && gt && out && workspace, "NULL pointer");
    T3D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && gt_h >= 1 && gt_w >= 1, "bad dims");
    T3D_REQUIRE(pred_stride >= 1 && pred_offset >= 0 && pred_offset < pred_stride, "bad pred stride/offset");
    T3D_REQUIRE((double)H * W < 1.0e9, "image too large");
    const int n = H * W, chunks = chunks_for(n);
    MetricsWs w = metrics_ws(workspace, B, n, chunks);
    if (workspace_bytes < w.total) { t3d_set_error("workspace too small"); return T3D_ERR_WORKSPACE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    T3D_CUDA(cudaMemsetAsync(w.counters, 0, w.zero_bytes, st));     // counters + select histograms
    dim3 g((unsigned)chunks, (unsigned)B);
    T3D_LAUNCH("depth_extract_kernel", st, depth_extract_kernel<<<g, kChunkThreads, 0, st>>>(
        pred, pred_stride, pred_offset, gt, gt_h, gt_w, mask, H, W, w.vz, w.vg, w.counters, w.hist));
    T3D_LAUNCH("select_pick_kernel", st, select_pick_kernel<0><<<B, 256, 0, st>>>(w.hist, w.state, w.counters, median_scaling, w.scale, out_medians));
    T3D_LAUNCH("select_hist_kernel", st, select_hist_kernel<1><<<g, kChunkThreads, 0, st>>>(w.vz, w.vg, w.state, w.counters, n, median_scaling, w.hist));
    T3D_LAUNCH("select_pick_kernel", st, select_pick_kernel<1><<<B, 256, 0, st>>>(w.hist, w.state, w.counters, median_scaling, w.scale, out_medians));
    T3D_LAUNCH("select_hist_kernel", st, select_hist_kernel<2><<<g, kChunkThreads, 0, st>>>(w.vz, w.vg, w.state, w.counters, n, median_scaling, w.hist));
    T3D_LAUNCH("select_pick_kernel", st, select_pick_kernel<2><<<B, 256, 0, st>>>(w.hist, w.state, w.counters, median_scaling, w.scale, out_medians));
    T3D_LAUNCH("metrics_sum_kernel", st, metrics_sum_kernel<<<g, kChunkThreads, 0, st>>>(w.vz, w.vg, w.scale, n, chunks, w.partials));
    T3D_LAUNCH("metrics_finalize_kernel", st, metrics_finalize_kernel<<<B, 32, 0, st>>>(w.partials, w.counters, chunks, out, out_f64));
    return T3D_OK;
}

int t3d_pointmap_to_depth(const float* pointmap, float* depth, size_t n_pixels, void* stream) {
    T3D_REQUIRE(pointmap && depth, "NULL pointer");
    if (n_pixels == 0) return T3D_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t blocks = (n_pixels + 255) / 256;
    const int grid = (int)(blocks < (size_t)t3d_sm_count() * 8 ? blocks : (size_t)t3d_sm_count() * 8);
    T3D_LAUNCH("pointmap_to_depth_kernel", st, pointmap_to_depth_kernel<<<grid, 256, 0, st>>>(pointmap, depth, n_pixels));
    return T3D_OK;
}

int t3d_estimate_focal(const float* pointmap, const float* depth, int B, int H, int W, double* out_K, void* stream) {
    T3D_REQUIRE(pointmap && out_K, "NULL pointer");
    T3D_REQUIRE(B >= 1 && H >= 1 && W >= 1, "bad dims");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    T3D_LAUNCH("focal_estimate_kernel", st, focal_estimate_kernel<<<B, t3d_select::kThreads, 0, st>>>(pointmap, depth, H, W, out_K));
    return T3D_OK;
}

int t3d_project_points(const float* pointmap, float fx, float fy, float cx, float cy, float* uv, size_t n_pixels,
                       void* stream) {
    T3D_REQUIRE(pointmap && uv, "NULL pointer");
    if (n_pixels == 0) return T3D_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t blocks = (n_pixels + 255) / 256;
    const int grid = (int)(blocks < (size_t)t3d_sm_count() * 8 ? blocks : (size_t)t3d_sm_count() * 8);
    T3D_LAUNCH("project_points_kernel", st, project_points_kernel<<<grid, 256, 0, st>>>(pointmap, fx, fy, cx, cy, uv, n_pixels));
    return T3D_OK;
}

}  // extern "C"
