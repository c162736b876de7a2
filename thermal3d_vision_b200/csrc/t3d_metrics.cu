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

struct MetricsWs {
    float* vz; float* vg; unsigned char* valid;
    int* counters;     // [B][4]: n_valid, pred_nan, gt_nan, pad
    float* scale;      // [B]
    double* partials;  // [B][chunks][kNPart]
    size_t total;
};

MetricsWs metrics_ws(void* base, int B, int n, int chunks) {
    MetricsWs w;
    size_t off = 0;
    char* p = reinterpret_cast<char*>(base);
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += t3d_align_up(bytes, 256); return r; };
    w.vz = reinterpret_cast<float*>(take((size_t)B * n * 4));
    w.vg = reinterpret_cast<float*>(take((size_t)B * n * 4));
    w.valid = reinterpret_cast<unsigned char*>(take((size_t)B * n));
    w.counters = reinterpret_cast<int*>(take((size_t)B * 4 * sizeof(int)));
    w.scale = reinterpret_cast<float*>(take((size_t)B * sizeof(float)));
    w.partials = reinterpret_cast<double*>(take((size_t)B * chunks * kNPart * sizeof(double)));
    w.total = off;
    return w;
}

int chunks_for(int n) { return max(1, min(96, (n + 4095) / 4096)); }

// ------------------------------------------------------------------ M1: extract (K5)
// pred element (b, i) lives at pred[(b*n + i) * pred_stride + pred_offset]: stride 3 / offset 2 reads
// the Z channel of an AoS pointmap in place (depth is never materialised by the caller).
__global__ void __launch_bounds__(kChunkThreads)
depth_extract_kernel(const float* __restrict__ pred, int pred_stride, int pred_offset,
                     const float* __restrict__ gt, int gt_h, int gt_w, const unsigned char* __restrict__ mask,
                     int H, int W, float* __restrict__ vz, float* __restrict__ vg,
                     unsigned char* __restrict__ valid, int* __restrict__ counters) {
    const int b = blockIdx.y, n = H * W;
    const bool resample = (gt_h != H) || (gt_w != W);
    const double fx = (double)gt_w / (double)W, fy = (double)gt_h / (double)H;
    const float* g = gt + (size_t)b * gt_h * gt_w;
    const float* p = pred + (size_t)b * n * pred_stride + pred_offset;
    int nv = 0, pnan = 0, gnan = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float gv;
        if (resample) {   // cv2 INTER_NEAREST (utils/evaluate_depth_metrics.py:321-323)
            const int y = i / W, x = i - y * W;
            const int sx = min((int)floor(__dmul_rn((double)x, fx)), gt_w - 1);
            const int sy = min((int)floor(__dmul_rn((double)y, fy)), gt_h - 1);
            gv = __ldg(g + (size_t)sy * gt_w + sx);
        } else {
            gv = __ldg(g + i);
        }
        const float pv = __ldg(p + (size_t)i * pred_stride);
        const bool ok = mask ? (mask[(size_t)b * n + i] != 0) : (gv > 0.f && isfinite(gv));   // utils/metrics.py:27
        vz[(size_t)b * n + i] = pv;
        vg[(size_t)b * n + i] = gv;
        valid[(size_t)b * n + i] = ok ? 1 : 0;
        if (ok) { ++nv; pnan += isnan(pv); gnan += isnan(gv); }
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    pnan = __reduce_add_sync(0xffffffffu, pnan);
    gnan = __reduce_add_sync(0xffffffffu, gnan);
    if ((threadIdx.x & 31) == 0) {
        if (nv) atomicAdd(&counters[4 * b], nv);
        if (pnan) atomicAdd(&counters[4 * b + 1], pnan);
        if (gnan) atomicAdd(&counters[4 * b + 2], gnan);
    }
}

// ------------------------------------------------------------------ M2: medians -> scale
// np.median of a float32 vector: mean of the two middle order statistics in fp32 (even n); NaN if any NaN.
__device__ float median_of(t3d_select::Smem& sm, const float* __restrict__ v, const unsigned char* __restrict__ valid,
                           int n, int n_valid, int n_nan) {
    if (n_nan > 0) return __int_as_float(0x7fc00000);
    auto get = [&](int i, float* out) { *out = v[i]; return valid[i] != 0; };
    const unsigned int r0 = (unsigned)(n_valid - 1) / 2, r1 = (unsigned)n_valid / 2;
    const float a = t3d_select::select_rank(sm, n, r0, get);
    if (r1 == r0) return a;
    const float b = t3d_select::select_rank(sm, n, r1, get);
    return __fmul_rn(__fadd_rn(a, b), 0.5f);
}

__global__ void __launch_bounds__(t3d_select::kThreads, 1)
median_scale_kernel(const float* __restrict__ vz, const float* __restrict__ vg, const unsigned char* __restrict__ valid,
                    const int* __restrict__ counters, int n, int median_scaling, float* __restrict__ scale,
                    float* __restrict__ out_medians) {
    __shared__ t3d_select::Smem sm;
    const int b = blockIdx.x;
    const int nv = counters[4 * b];
    float s = 1.0f, mg = 0.f, mp = 0.f;
    if (nv > 0 && median_scaling) {
        mg = median_of(sm, vg + (size_t)b * n, valid + (size_t)b * n, n, nv, counters[4 * b + 2]);
        mp = median_of(sm, vz + (size_t)b * n, valid + (size_t)b * n, n, nv, counters[4 * b + 1]);
        s = __fdiv_rn(mg, mp);                                   // utils/metrics.py:47
    }
    if (threadIdx.x == 0) {
        scale[b] = s;
        if (out_medians) { out_medians[2 * b] = mg; out_medians[2 * b + 1] = mp; }
    }
}

// ------------------------------------------------------------------ M3: per-pixel terms
__device__ __forceinline__ float np_maximum(float a, float b) { return (isnan(a) || isnan(b)) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

__global__ void __launch_bounds__(kChunkThreads)
metrics_sum_kernel(const float* __restrict__ vz, const float* __restrict__ vg, const unsigned char* __restrict__ valid,
                   const float* __restrict__ scale, int n, int chunks, double* __restrict__ partials) {
    __shared__ double red[kChunkThreads / 32][kNPart];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const float s = scale[b];
    const int per = (n + chunks - 1) / chunks;
    const int i0 = chunk * per, i1 = min(i0 + per, n);
    double acc[4] = {0, 0, 0, 0};
    int cnt[3] = {0, 0, 0};
    for (int i = i0 + threadIdx.x; i < i1; i += kChunkThreads) {
        if (!valid[(size_t)b * n + i]) continue;
        const float gt = vg[(size_t)b * n + i];
        const float pr = __fmul_rn(vz[(size_t)b * n + i], s);                    // pred *= scale   (:48)
        const float th = np_maximum(__fdiv_rn(gt, pr), __fdiv_rn(pr, gt));       // :51
        cnt[0] += th < 1.25f; cnt[1] += th < 1.5625f; cnt[2] += th < 1.953125f;  // :52-54
        const float d = __fsub_rn(gt, pr);
        const float d2 = __fmul_rn(d, d);
        acc[0] += (double)__fdiv_rn(fabsf(d), gt);                               // :56
        acc[1] += (double)__fdiv_rn(d2, gt);                                     // :57
        acc[2] += (double)d2;                                                    // :58
        const float dl = __fsub_rn(logf(gt), logf(pr));
        acc[3] += (double)__fmul_rn(dl, dl);                                     // :59
    }
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
    T3D_CUDA(cudaMemsetAsync(w.counters, 0, (size_t)B * 4 * sizeof(int), st));
    dim3 g((unsigned)chunks, (unsigned)B);
    T3D_LAUNCH("depth_extract_kernel", st, depth_extract_kernel<<<g, kChunkThreads, 0, st>>>(pred, pred_stride, pred_offset, gt, gt_h, gt_w, mask, H, W,
                                                      w.vz, w.vg, w.valid, w.counters));
    T3D_LAUNCH("median_scale_kernel", st, median_scale_kernel<<<B, t3d_select::kThreads, 0, st>>>(w.vz, w.vg, w.valid, w.counters, n, median_scaling,
                                                            w.scale, out_medians));
    T3D_LAUNCH("metrics_sum_kernel", st, metrics_sum_kernel<<<g, kChunkThreads, 0, st>>>(w.vz, w.vg, w.valid, w.scale, n, chunks, w.partials));
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
