// t3d_fire.cu -- image operators of the experimental fire-scene pipeline
// (/root/reference/thermal_dustr_inference_for_experiment.py:62-377), sm_100a.
//
// The reference composes these from OpenCV / NumPy / SciPy calls; each kernel restates the library's published
// algorithm so that integer / byte results are bit-identical (CLAHE, Canny, histogram) and float results agree to
// rounding (Sobel, bilateral):
//   CLAHE      cv2.createCLAHE(clipLimit, (8,8)).apply(u8)  (:108-109, :220-221)   modules/imgproc/src/clahe.cpp
//   Canny      cv2.Canny(u8, low, high)                     (:135, :225)           modules/imgproc/src/canny.cpp
//   Sobel      cv2.Sobel(f32, CV_32F, 1|0, 0|1, ksize=3)    (:228-229)
//   bilateral  cv2.bilateralFilter(f32, d, sc, ss)          (:273, :375)           bilateral_filter.simd.hpp
//   histogram  np.histogram(x, bins=100, range=(0, 1))      (:188)
//   outliers   3-sigma mask + 5x5 median of the inliers     (:335-356)
// plus the per-pixel compositions of preprocess_fire_scene_thermal (:62-152) and
// advanced_fire_scene_processing (:154-282).  Not a hot path: simple one-thread-per-pixel kernels.
#include "t3d_common.cuh"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {          // BORDER_REFLECT_101: -1 -> 1, n -> n - 2
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}
__device__ __forceinline__ int clampi(int i, int n) { return min(max(i, 0), n - 1); }    // BORDER_REPLICATE
__device__ __forceinline__ unsigned char sat_u8(float v) {          // saturate_cast<uchar>(float): cvRound, clamp
    const int r = __float2int_rn(v);
    return (unsigned char)min(max(r, 0), 255);
}

// ------------------------------------------------------------------ CLAHE (clahe.cpp)
// grid (tiles, B); the tile grid covers the image extended (reflect 101) at the bottom / right to a multiple of it
__global__ void __launch_bounds__(256) clahe_lut_kernel(const unsigned char* __restrict__ src, int H, int W, int tiles_x,
                                                        int th, int tw, int clip, float lut_scale, unsigned char* __restrict__ luts) {
    __shared__ int hist[256];
    __shared__ int wsum[8];
    __shared__ int s_clipped;
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const unsigned char* img = src + (size_t)b * H * W;
    hist[tid] = 0;
    if (tid == 0) s_clipped = 0;
    __syncthreads();
    for (int i = tid; i < th * tw; i += 256) {
        const int r = i / tw, c = i - r * tw;
        const int y = reflect101(ty * th + r, H), x = reflect101(tx * tw + c, W);
        atomicAdd(&hist[img[(size_t)y * W + x]], 1);
    }
    __syncthreads();
    int v = hist[tid];
    if (clip > 0) {
        const int excess = max(v - clip, 0);
        const int e = __reduce_add_sync(0xffffffffu, excess);
        if (lane == 0 && e) atomicAdd(&s_clipped, e);
        __syncthreads();
        const int clipped = s_clipped;
        v = min(v, clip);
        const int batch = clipped / 256;
        int residual = clipped - batch * 256;
        v += batch;
        if (residual != 0) {                                    // for (i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++
            const int step = max(256 / residual, 1);
            if (tid % step == 0 && tid / step < residual) v += 1;
        }
    }
    // inclusive prefix sum over the 256 bins
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wrp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < wrp; ++w) before += wsum[w];
    const int sum = before + incl;
    luts[((size_t)b * gridDim.x + tile) * 256 + tid] = sat_u8(__fmul_rn((float)sum, lut_scale));
}

__global__ void __launch_bounds__(256) clahe_interp_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                                           const unsigned char* __restrict__ luts, int H, int W,
                                                           int tiles_x, int tiles_y, float inv_tw, float inv_th) {
    const int b = blockIdx.y;
    const size_t n = (size_t)H * W;
    const unsigned char* lut = luts + (size_t)b * tiles_x * tiles_y * 256;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f), tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
        int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
        const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
        const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
        const int tx2 = min(tx1 + 1, tiles_x - 1), ty2 = min(ty1 + 1, tiles_y - 1);
        tx1 = max(tx1, 0); ty1 = max(ty1, 0);
        const int v = src[(size_t)b * n + i];
        const float l11 = (float)lut[(ty1 * tiles_x + tx1) * 256 + v], l12 = (float)lut[(ty1 * tiles_x + tx2) * 256 + v];
        const float l21 = (float)lut[(ty2 * tiles_x + tx1) * 256 + v], l22 = (float)lut[(ty2 * tiles_x + tx2) * 256 + v];
        const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
        const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
        dst[(size_t)b * n + i] = sat_u8(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
    }
}

// ------------------------------------------------------------------ Canny (canny.cpp; aperture 3, L1 magnitude)
// Sobel with BORDER_REPLICATE (what cv::Canny asks for), |dx| + |dy|
__global__ void __launch_bounds__(256) canny_grad_kernel(const unsigned char* __restrict__ src, int H, int W,
                                                         short* __restrict__ dx, short* __restrict__ dy, int* __restrict__ mag) {
    const size_t n = (size_t)H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        auto P = [&](int yy, int xx) { return (int)src[(size_t)clampi(yy, H) * W + clampi(xx, W)]; };
        const int gx = (P(y - 1, x + 1) - P(y - 1, x - 1)) + 2 * (P(y, x + 1) - P(y, x - 1)) + (P(y + 1, x + 1) - P(y + 1, x - 1));
        const int gy = (P(y + 1, x - 1) + 2 * P(y + 1, x) + P(y + 1, x + 1)) - (P(y - 1, x - 1) + 2 * P(y - 1, x) + P(y - 1, x + 1));
        dx[i] = (short)gx; dy[i] = (short)gy; mag[i] = abs(gx) + abs(gy);
    }
}
// non-maximum suppression with the fixed-point tangent tests; map: 2 = edge seed (m > high), 0 = candidate, 1 = no edge
__global__ void __launch_bounds__(256) canny_nms_kernel(const short* __restrict__ dx, const short* __restrict__ dy,
                                                        const int* __restrict__ mag, int H, int W, int low, int high,
                                                        unsigned char* __restrict__ map) {
    const size_t n = (size_t)H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        auto M = [&](int yy, int xx) { return (yy < 0 || yy >= H || xx < 0 || xx >= W) ? 0 : mag[(size_t)yy * W + xx]; };   // zero border
        const int m = mag[i];
        unsigned char r = 1;
        if (m > low) {
            const int xs = dx[i], ys = dy[i];
            const long long ax = abs(xs), ay = (long long)abs(ys) << 15;
            const long long tg22 = ax * 13573;                   // tan(22.5 deg) * 2^15
            bool keep;
            if (ay < tg22) keep = m > M(y, x - 1) && m >= M(y, x + 1);
            else {
                const long long tg67 = tg22 + (ax << 16);
                if (ay > tg67) keep = m > M(y - 1, x) && m >= M(y + 1, x);
                else { const int s = ((xs ^ ys) < 0) ? -1 : 1; keep = m > M(y - 1, x - s) && m > M(y + 1, x + s); }
            }
            if (keep) r = (m > high) ? 2 : 0;
        }
        map[i] = r;
    }
}
// hysteresis: a candidate 8-connected to an edge becomes an edge.  One launch = each 32x32 tile (+ halo) relaxed to its
// fixed point in shared memory; the host repeats launches until no tile changed.
__global__ void __launch_bounds__(1024) canny_hyst_kernel(unsigned char* __restrict__ map, int H, int W, int* __restrict__ changed) {
    __shared__ unsigned char t[34][34];
    __shared__ int s_any, s_iter;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int k = threadIdx.x; k < 34 * 34; k += 1024) {
        const int r = k / 34, c = k - r * 34, y = y0 + r - 1, x = x0 + c - 1;
        t[r][c] = (y >= 0 && y < H && x >= 0 && x < W) ? map[(size_t)y * W + x] : (unsigned char)1;
    }
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    const int y = y0 + ty, x = x0 + tx;
    const bool inside = y < H && x < W;
    bool mine_changed = false;
    for (;;) {
        if (threadIdx.x == 0) s_iter = 0;
        __syncthreads();
        if (inside && t[ty + 1][tx + 1] == 0) {
            const bool nb = t[ty][tx] == 2 || t[ty][tx + 1] == 2 || t[ty][tx + 2] == 2 || t[ty + 1][tx] == 2 ||
                            t[ty + 1][tx + 2] == 2 || t[ty + 2][tx] == 2 || t[ty + 2][tx + 1] == 2 || t[ty + 2][tx + 2] == 2;
            if (nb) { t[ty + 1][tx + 1] = 2; mine_changed = true; s_iter = 1; }
        }
        __syncthreads();
        if (!s_iter) break;
        __syncthreads();
    }
    if (mine_changed) { map[(size_t)y * W + x] = 2; s_any = 1; }
    __syncthreads();
    if (threadIdx.x == 0 && s_any) atomicExch(changed, 1);
}
__global__ void __launch_bounds__(256) canny_final_kernel(const unsigned char* __restrict__ map, unsigned char* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (map[i] == 2) ? 255 : 0;
}

// ------------------------------------------------------------------ Sobel 3x3 on float (BORDER_REFLECT_101)
__global__ void __launch_bounds__(256) sobel3_f32_kernel(const float* __restrict__ src, float* __restrict__ dx, float* __restrict__ dy,
                                                         int H, int W) {
    const size_t n = (size_t)H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H), xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
        auto P = [&](int yy, int xx) { return src[(size_t)yy * W + xx]; };
        const float r0 = P(ym, xp) - P(ym, xm), r1 = P(y, xp) - P(y, xm), r2 = P(yp, xp) - P(yp, xm);
        dx[i] = __fadd_rn(__fmul_rn(r1, 2.0f), __fadd_rn(r0, r2));
        const float s0 = __fadd_rn(__fmul_rn(P(ym, x), 2.0f), __fadd_rn(P(ym, xm), P(ym, xp)));
        const float s2 = __fadd_rn(__fmul_rn(P(yp, x), 2.0f), __fadd_rn(P(yp, xm), P(yp, xp)));
        dy[i] = s2 - s0;
    }
}

// ------------------------------------------------------------------ np.histogram(x, bins=100, range=(0, 1))
// numpy: keep first <= x <= last; idx = int((x - first) / (last - first) * 100) in fp64; idx == 100 -> 99; then the
// two corrections against the fp64 edges linspace(0, 1, 101) (x < edge[idx] -> idx - 1; x >= edge[idx + 1] -> idx + 1).
__global__ void __launch_bounds__(256) hist100_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[100];
    if (threadIdx.x < 100) sh[threadIdx.x] = 0u;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double a = (double)x[i];
        if (!(a >= 0.0 && a <= 1.0)) continue;
        int idx = (int)(a / 1.0 * 100.0);
        if (idx == 100) idx = 99;
        auto edge = [](int k) { return (k == 100) ? 1.0 : (double)k * 0.01; };      // np.linspace(0, 1, 101): k * step, last = stop
        if (a < edge(idx)) idx -= 1;
        else if (a >= edge(idx + 1) && idx != 99) idx += 1;
        atomicAdd(&sh[idx], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 100 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// ------------------------------------------------------------------ bilateral filter on float, channels last (C = 1 or 3)
// cv2: circular window of radius d / 2, BORDER_REFLECT_101, space weight exp(-r^2 / (2 ss^2)), colour weight
// exp(-(sum_c |v - v0|)^2 / (2 sc^2)) (tabulated and interpolated in OpenCV: equal to rounding), centre weight 1;
// an image whose value range is below FLT_EPSILON is copied.
template <int C>
__global__ void __launch_bounds__(256) bilateral_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W,
                                                            int radius, float gauss_color, float gauss_space,
                                                            const float* __restrict__ minmax) {
    const size_t n = (size_t)H * W;
    const bool flat = fabsf(minmax[1] - minmax[0]) < 1.1920929e-07f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        float v0[C], sum[C], wsum = 1.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) { v0[c] = src[i * C + c]; sum[c] = v0[c]; }
        if (!flat) {
            for (int dyy = -radius; dyy <= radius; ++dyy) {
                for (int dxx = -radius; dxx <= radius; ++dxx) {
                    const int r2 = dyy * dyy + dxx * dxx;
                    if (r2 == 0 || r2 > radius * radius) continue;
                    const size_t j = (size_t)reflect101(y + dyy, H) * W + reflect101(x + dxx, W);
                    float v[C], diff = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) { v[c] = src[j * C + c]; diff += fabsf(v[c] - v0[c]); }
                    const float w = (float)exp((double)r2 * (double)gauss_space) * (float)exp((double)diff * (double)diff * (double)gauss_color);
                    wsum += w;
#pragma unroll
                    for (int c = 0; c < C; ++c) sum[c] = fmaf(v[c], w, sum[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) dst[i * C + c] = flat ? v0[c] : sum[c] / wsum;
    }
}

// min / max / nan-aware sum and sum of squares of a float array: out[0] min, [1] max (as floats) -- one CTA, fixed order
__global__ void __launch_bounds__(1024) minmax_kernel(const float* __restrict__ x, size_t n, float* __restrict__ out) {
    __shared__ float smin[32], smax[32];
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    for (size_t i = threadIdx.x; i < n; i += 1024) { const float v = x[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    mn = warp_min(mn); mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        mn = warp_min(smin[threadIdx.x]); mx = warp_max(smax[threadIdx.x]);
        if (threadIdx.x == 0) { out[0] = mn; out[1] = mx; }
    }
}

// ------------------------------------------------------------------ depth refinement (:335-356)
// stats[0] = nanmean, stats[1] = nanstd (population), as float32 like numpy's results for a float32 array
__global__ void __launch_bounds__(1024) nanstats_kernel(const float* __restrict__ x, size_t n, float* __restrict__ stats) {
    __shared__ double sa[32], sb[32];
    __shared__ double s_mean, s_cnt;
    double s = 0.0, c = 0.0;
    for (size_t i = threadIdx.x; i < n; i += 1024) { const float v = x[i]; if (v == v) { s += (double)v; c += 1.0; } }
    s = warp_sum(s); c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = s; sb[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x < 32) {
        s = warp_sum(sa[threadIdx.x]); c = warp_sum(sb[threadIdx.x]);
        if (threadIdx.x == 0) { s_cnt = c; s_mean = (c > 0.0) ? s / c : __longlong_as_double(0x7ff8000000000000LL); }
    }
    __syncthreads();
    const double mean = s_mean;
    double q = 0.0;
    for (size_t i = threadIdx.x; i < n; i += 1024) { const float v = x[i]; if (v == v) { const double d = (double)v - mean; q += d * d; } }
    q = warp_sum(q);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sa[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x < 32) {
        q = warp_sum(sa[threadIdx.x]);
        if (threadIdx.x == 0) { stats[0] = (float)mean; stats[1] = (s_cnt > 0.0) ? (float)sqrt(q / s_cnt) : __int_as_float(0x7fc00000); }
    }
}
// outliers (|d - mean| > 3 std, float32 arithmetic) -> median of the 5x5 neighbours that are not outliers (np.median:
// fp32 mean of the two middle values for an even count); no inlier in the window -> the mean
__global__ void __launch_bounds__(256) outlier_median_kernel(const float* __restrict__ d, float* __restrict__ out, int H, int W,
                                                             const float* __restrict__ stats, unsigned char* __restrict__ mask_out) {
    const size_t n = (size_t)H * W;
    const float mean = stats[0], thr = __fmul_rn(3.0f, stats[1]);
    auto outlier = [&](float v) { return fabsf(__fsub_rn(v, mean)) > thr; };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = d[i];
        const bool o = outlier(v);
        if (mask_out) mask_out[i] = o ? 1 : 0;
        if (!o) { out[i] = v; continue; }
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        float buf[25];
        int m = 0;
        bool has_nan = false;
        for (int yy = max(0, y - 2); yy < min(H, y + 3); ++yy)
            for (int xx = max(0, x - 2); xx < min(W, x + 3); ++xx) {
                const float u = d[(size_t)yy * W + xx];
                if (outlier(u)) continue;
                if (u != u) has_nan = true;
                int k = m++;                                      // insertion sort
                while (k > 0 && buf[k - 1] > u) { buf[k] = buf[k - 1]; --k; }
                buf[k] = u;
            }
        float r;
        if (m == 0) r = mean;
        else if (has_nan) r = __int_as_float(0x7fc00000);
        else r = (m & 1) ? buf[m >> 1] : __fmul_rn(__fadd_rn(buf[(m >> 1) - 1], buf[m >> 1]), 0.5f);
        out[i] = r;
    }
}

// ------------------------------------------------------------------ per-pixel compositions
// gray = 0.299 c0 + 0.587 c1 + 0.114 c2 of a [3,H,W] float image (fp32, left to right, no contraction), or plane 0
__global__ void __launch_bounds__(256) fire_gray_kernel(const float* __restrict__ img, int channels, size_t n, float* __restrict__ gray) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        gray[i] = (channels >= 3) ? gray3(img[i], img[n + i], img[2 * n + i]) : img[i];
}
// preprocess_fire_scene_thermal, steps 1-3 (:95-107,135): thermal_norm (fp64) -> the two uint8 images CLAHE and Canny take
__global__ void __launch_bounds__(256) fire_norm_u8_kernel(const float* __restrict__ gray, size_t n, const double* __restrict__ pct,
                                                           unsigned char* __restrict__ base_u8, unsigned char* __restrict__ norm_u8) {
    const double lo = pct[0], hi = pct[1], den = (hi - lo) + 1e-6;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double g = (double)gray[i];
        const double tn = (fmin(fmax(g, lo), hi) - lo) / den;                        // np.clip propagates NaN the same way
        const double base = fmin(fmax((1.0 - tn) * 1.2, 0.0), 1.0);
        base_u8[i] = (unsigned char)(int)(base * 255.0);                             // astype(uint8): truncation
        norm_u8[i] = (unsigned char)(int)(tn * 255.0);
    }
}
// preprocess_fire_scene_thermal, steps 3-7 (:100-146); noise = np.random.rand(h, w).astype(float32) drawn by the host
__global__ void __launch_bounds__(256) fire_compose_kernel(const float* __restrict__ gray, size_t n, const double* __restrict__ pct,
                                                           const unsigned char* __restrict__ clahe, const unsigned char* __restrict__ edges,
                                                           const float* __restrict__ noise, float* __restrict__ out) {
    const double lo = pct[0], hi = pct[1], den = (hi - lo) + 1e-6;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double tn = (fmin(fmax((double)gray[i], lo), hi) - lo) / den;
        const bool fire = tn > 0.7;
        const float base = __fdiv_rn((float)clahe[i], 255.0f);
        const float nz = noise ? __fmul_rn(noise[i], 0.1f) : 0.f;
        const float e = __fdiv_rn((float)edges[i], 255.0f);
        const double ew = fire ? 0.3 : 0.15;
        const float fire_rgb[3] = {0.8f, 0.3f, 0.1f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float r = fire ? __fadd_rn(fire_rgb[c], nz) : base;
            r = (float)__dadd_rn(__dmul_rn((double)r, 1.0 - ew), __dmul_rn((double)e, ew));   // float32 * float64 array -> float64 -> stored float32
            out[(size_t)c * n + i] = fminf(fmaxf(r, 0.f), 1.f);
        }
    }
}
// advanced_fire_scene_processing: the two uint8 images (:219, :225)
__global__ void __launch_bounds__(256) adv_u8_kernel(const float* __restrict__ gray, size_t n, unsigned char* __restrict__ inv_u8,
                                                     unsigned char* __restrict__ gray_u8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float g = gray[i];
        inv_u8[i] = (unsigned char)(int)__fmul_rn(__fsub_rn(1.0f, g), 255.0f);
        gray_u8[i] = (unsigned char)(int)__fmul_rn(g, 255.0f);
    }
}
__global__ void __launch_bounds__(256) sobel_mag_kernel(const float* __restrict__ dx, const float* __restrict__ dy, size_t n, float* __restrict__ mag) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        mag[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx[i], dx[i]), __fmul_rn(dy[i], dy[i])));
}
// advanced_fire_scene_processing, steps 5-8 (:231-270) -> channels-LAST [H,W,3] float image for the bilateral filter
__global__ void __launch_bounds__(256) adv_compose_kernel(const float* __restrict__ gray, size_t n, double fire_threshold,
                                                          const unsigned char* __restrict__ clahe, const unsigned char* __restrict__ canny,
                                                          const float* __restrict__ mag, const float* __restrict__ mag_minmax,
                                                          const float* __restrict__ noise, float* __restrict__ out_hwc) {
    const float mn = mag_minmax[0], den = __fadd_rn(__fsub_rn(mag_minmax[1], mn), 1e-6f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const bool fire = (double)gray[i] > fire_threshold;
        const float cl = __fdiv_rn((float)clahe[i], 255.0f);
        const float e1 = __fdiv_rn((float)canny[i], 255.0f);
        const float sm = __fdiv_rn(__fsub_rn(mag[i], mn), den);
        const float edge = fmaxf(e1, sm);
        const float nz = noise ? __fmul_rn(noise[i], 0.15f) : 0.f;
        const float es = fire ? 0.4f : 0.2f;
        const float k[3] = {0.5f, 0.3f, 0.2f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float r = fire ? __fadd_rn(__fmul_rn(cl, k[c]), nz) : cl;
            out_hwc[i * 3 + c] = __fadd_rn(__fmul_rn(r, __fsub_rn(1.0f, es)), __fmul_rn(edge, es));
        }
    }
}
// clip(0, 1) + channels-last [H,W,3] -> planar [3,H,W]
__global__ void __launch_bounds__(256) hwc_to_chw_clip_kernel(const float* __restrict__ hwc, size_t n, float* __restrict__ chw) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) chw[(size_t)c * n + i] = fminf(fmaxf(hwc[i * 3 + c], 0.f), 1.f);
}

int grid_for(size_t n) {
    const size_t need = (n + 255) / 256, cap = (size_t)t3d_sm_count() * 8;
    return (int)(need < cap ? (need ? need : 1) : cap);
}

struct ClaheGeom { int tiles_x, tiles_y, th, tw, clip; float lut_scale, inv_tw, inv_th; };
ClaheGeom clahe_geom(int H, int W, double clip_limit, int tiles_x, int tiles_y) {
    ClaheGeom g;
    g.tiles_x = tiles_x; g.tiles_y = tiles_y;
    const int eh = (H % tiles_y == 0) ? H : H + (tiles_y - H % tiles_y), ew = (W % tiles_x == 0) ? W : W + (tiles_x - W % tiles_x);
    g.th = eh / tiles_y; g.tw = ew / tiles_x;
    const int area = g.th * g.tw;
    g.lut_scale = 255.0f / (float)area;
    g.clip = 0;
    if (clip_limit > 0.0) { g.clip = (int)(clip_limit * area / 256); if (g.clip < 1) g.clip = 1; }
    g.inv_tw = 1.0f / (float)g.tw; g.inv_th = 1.0f / (float)g.th;
    return g;
}

}  // namespace

extern "C" {

size_t t3d_clahe_workspace_bytes(int B, int tiles_x, int tiles_y) {
    return (B < 1 || tiles_x < 1 || tiles_y < 1) ? 0 : t3d_align_up((size_t)B * tiles_x * tiles_y * 256, 256);
}

int t3d_clahe_u8(const unsigned char* src, unsigned char* dst, int B, int H, int W, double clip_limit, int tiles_x, int tiles_y,
                 void* workspace, size_t workspace_bytes, void* stream) {
    T3D_REQUIRE(src && dst && workspace, "NULL pointer");
    T3D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && tiles_x >= 1 && tiles_y >= 1 && tiles_x * tiles_y <= 65535, "bad dims");
    // OpenCV extends the image by tiles - (size % tiles) rows / columns with BORDER_REFLECT_101: needs size > that
    T3D_REQUIRE(H > tiles_y && W > tiles_x, "image smaller than the tile grid");
    if (workspace_bytes < t3d_clahe_workspace_bytes(B, tiles_x, tiles_y)) { t3d_set_error("workspace too small"); return T3D_ERR_WORKSPACE; }
    const ClaheGeom g = clahe_geom(H, W, clip_limit, tiles_x, tiles_y);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* luts = reinterpret_cast<unsigned char*>(workspace);
    T3D_LAUNCH("clahe_lut_kernel", st, clahe_lut_kernel<<<dim3(tiles_x * tiles_y, B), 256, 0, st>>>(src, H, W, tiles_x, g.th, g.tw, g.clip, g.lut_scale, luts));
    T3D_LAUNCH("clahe_interp_kernel", st, clahe_interp_kernel<<<dim3(grid_for((size_t)H * W), B), 256, 0, st>>>(
        src, dst, luts, H, W, tiles_x, tiles_y, g.inv_tw, g.inv_th));
    return T3D_OK;
}

size_t t3d_canny_workspace_bytes(int H, int W) {
    if (H < 1 || W < 1) return 0;
    const size_t n = (size_t)H * W;
    return t3d_align_up(n * 2, 256) * 2 + t3d_align_up(n * 4, 256) + t3d_align_up(n, 256) + 256;
}

/* cv2.Canny(src, low, high) (aperture 3, L1 gradient).  Synchronises `stream`: the hysteresis is iterated until no
 * tile changes and the host reads that flag. */
int t3d_canny_u8(const unsigned char* src, unsigned char* dst, int H, int W, double low_thresh, double high_thresh,
                 void* workspace, size_t workspace_bytes, void* stream) {
    T3D_REQUIRE(src && dst && workspace, "NULL pointer");
    T3D_REQUIRE(H >= 1 && W >= 1, "bad dims");
    if (workspace_bytes < t3d_canny_workspace_bytes(H, W)) { t3d_set_error("workspace too small"); return T3D_ERR_WORKSPACE; }
    const size_t n = (size_t)H * W;
    char* p = reinterpret_cast<char*>(workspace);
    short* dx = reinterpret_cast<short*>(p); p += t3d_align_up(n * 2, 256);
    short* dy = reinterpret_cast<short*>(p); p += t3d_align_up(n * 2, 256);
    int* mag = reinterpret_cast<int*>(p); p += t3d_align_up(n * 4, 256);
    unsigned char* map = reinterpret_cast<unsigned char*>(p); p += t3d_align_up(n, 256);
    int* changed = reinterpret_cast<int*>(p);
    if (low_thresh > high_thresh) { const double t = low_thresh; low_thresh = high_thresh; high_thresh = t; }
    const int low = (int)floor(low_thresh), high = (int)floor(high_thresh);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int g = grid_for(n);
    T3D_LAUNCH("canny_grad_kernel", st, canny_grad_kernel<<<g, 256, 0, st>>>(src, H, W, dx, dy, mag));
    T3D_LAUNCH("canny_nms_kernel", st, canny_nms_kernel<<<g, 256, 0, st>>>(dx, dy, mag, H, W, low, high, map));
    const dim3 tiles((W + 31) / 32, (H + 31) / 32);
    for (int round = 0; round < 100000; ++round) {
        T3D_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), st));
        for (int k = 0; k < 4; ++k)
            T3D_LAUNCH("canny_hyst_kernel", st, canny_hyst_kernel<<<tiles, 1024, 0, st>>>(map, H, W, changed));
        int host_changed = 0;
        T3D_CUDA(cudaMemcpyAsync(&host_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
        T3D_CUDA(cudaStreamSynchronize(st));
        if (!host_changed) break;
    }
    T3D_LAUNCH("canny_final_kernel", st, canny_final_kernel<<<g, 256, 0, st>>>(map, dst, n));
    return T3D_OK;
}

int t3d_sobel3_f32(const float* src, float* dx, float* dy, int H, int W, void* stream) {
    T3D_REQUIRE(src && dx && dy, "NULL pointer");
    T3D_REQUIRE(H >= 1 && W >= 1, "bad dims");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    T3D_LAUNCH("sobel3_f32_kernel", st, sobel3_f32_kernel<<<grid_for((size_t)H * W), 256, 0, st>>>(src, dx, dy, H, W));
    return T3D_OK;
}

int t3d_histogram100(const float* x, size_t n, unsigned int* hist100, void* stream) {
    T3D_REQUIRE(x && hist100, "NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    T3D_CUDA(cudaMemsetAsync(hist100, 0, 100 * sizeof(unsigned int), st));
    if (n == 0) return T3D_OK;
    T3D_LAUNCH("hist100_kernel", st, hist100_kernel<<<grid_for(n), 256, 0, st>>>(x, n, hist100));
    return T3D_OK;
}

/* cv2.bilateralFilter(src, d, sigma_color, sigma_space) for float32 images, channels last (C = 1 or 3).
 * scratch: 2 floats (the image's min / max). */
int t3d_bilateral_f32(const float* src, float* dst, int H, int W, int channels, int d, double sigma_color, double sigma_space,
                      float* scratch2, void* stream) {
    T3D_REQUIRE(src && dst && scratch2 && src != dst, "NULL / aliased pointer");
    T3D_REQUIRE(H >= 1 && W >= 1 && (channels == 1 || channels == 3), "bad dims (channels must be 1 or 3)");
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    int radius = (d <= 0) ? (int)lrint(sigma_space * 1.5) : d / 2;
    radius = radius < 1 ? 1 : radius;
    const float gc = (float)(-0.5 / (sigma_color * sigma_color)), gs = (float)(-0.5 / (sigma_space * sigma_space));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)H * W;
    T3D_LAUNCH("minmax_kernel", st, minmax_kernel<<<1, 1024, 0, st>>>(src, n * channels, scratch2));
    if (channels == 1) T3D_LAUNCH("bilateral_f32_kernel", st, bilateral_f32_kernel<1><<<grid_for(n), 256, 0, st>>>(src, dst, H, W, radius, gc, gs, scratch2));
    else T3D_LAUNCH("bilateral_f32_kernel", st, bilateral_f32_kernel<3><<<grid_for(n), 256, 0, st>>>(src, dst, H, W, radius, gc, gs, scratch2));
    return T3D_OK;
}

/* Step 1 of depth_refinement_with_outlier_removal (:335-356): stats2 <- {nanmean, nanstd}; out <- depth with every
 * 3-sigma outlier replaced by the median of its 5x5 inlier neighbours; mask (nullable) <- the outlier mask. */
int t3d_depth_outlier_median(const float* depth, float* out, int H, int W, float* stats2, unsigned char* mask, void* stream) {
    T3D_REQUIRE(depth && out && stats2 && depth != out, "NULL / aliased pointer");
    T3D_REQUIRE(H >= 1 && W >= 1, "bad dims");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)H * W;
    T3D_LAUNCH("nanstats_kernel", st, nanstats_kernel<<<1, 1024, 0, st>>>(depth, n, stats2));
    T3D_LAUNCH("outlier_median_kernel", st, outlier_median_kernel<<<grid_for(n), 256, 0, st>>>(depth, out, H, W, stats2, mask));
    return T3D_OK;
}

int t3d_fire_gray(const float* img_chw, int channels, int H, int W, float* gray, void* stream) {
    T3D_REQUIRE(img_chw && gray && channels >= 1, "bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)H * W;
    T3D_LAUNCH("fire_gray_kernel", st, fire_gray_kernel<<<grid_for(n), 256, 0, st>>>(img_chw, channels, n, gray));
    return T3D_OK;
}

int t3d_fire_norm_u8(const float* gray, int H, int W, const double* percentiles2, unsigned char* base_u8, unsigned char* norm_u8, void* stream) {
    T3D_REQUIRE(gray && percentiles2 && base_u8 && norm_u8, "NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)H * W;
    T3D_LAUNCH("fire_norm_u8_kernel", st, fire_norm_u8_kernel<<<grid_for(n), 256, 0, st>>>(gray, n, percentiles2, base_u8, norm_u8));
    return T3D_OK;
}

int t3d_fire_compose(const float* gray, int H, int W, const double* percentiles2, const unsigned char* clahe_u8,
                     const unsigned char* canny_u8, const float* noise, float* out_chw, void* stream) {
    T3D_REQUIRE(gray && percentiles2 && clahe_u8 && canny_u8 && out_chw, "NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)H * W;
    T3D_LAUNCH("fire_compose_kernel", st, fire_compose_kernel<<<grid_for(n), 256, 0, st>>>(gray, n, percentiles2, clahe_u8, canny_u8, noise, out_chw));
    return T3D_OK;
}

int t3d_fire_adv_u8(const float* gray, int H, int W, unsigned char* inverted_u8, unsigned char* gray_u8, void* stream) {
    T3D_REQUIRE(gray && inverted_u8 && gray_u8, "NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)H * W;
    T3D_LAUNCH("adv_u8_kernel", st, adv_u8_kernel<<<grid_for(n), 256, 0, st>>>(gray, n, inverted_u8, gray_u8));
    return T3D_OK;
}

/* steps 5-8 of advanced_fire_scene_processing; scratch: 2 * H * W + 2 floats (Sobel x / y -> magnitude, its min / max).
 * out_hwc [H,W,3] is the input of the 9/75/75 bilateral filter. */
int t3d_fire_adv_compose(const float* gray, int H, int W, double fire_threshold, const unsigned char* clahe_u8,
                         const unsigned char* canny_u8, const float* noise, float* out_hwc, float* scratch, void* stream) {
    T3D_REQUIRE(gray && clahe_u8 && canny_u8 && out_hwc && scratch, "NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)H * W;
    float* dx = scratch; float* dy = scratch + n; float* mm = scratch + 2 * n;
    T3D_LAUNCH("sobel3_f32_kernel", st, sobel3_f32_kernel<<<grid_for(n), 256, 0, st>>>(gray, dx, dy, H, W));
    T3D_LAUNCH("sobel_mag_kernel", st, sobel_mag_kernel<<<grid_for(n), 256, 0, st>>>(dx, dy, n, dx));
    T3D_LAUNCH("minmax_kernel", st, minmax_kernel<<<1, 1024, 0, st>>>(dx, n, mm));
    T3D_LAUNCH("adv_compose_kernel", st, adv_compose_kernel<<<grid_for(n), 256, 0, st>>>(gray, n, fire_threshold, clahe_u8, canny_u8, dx, mm, noise, out_hwc));
    return T3D_OK;
}

int t3d_hwc_to_chw_clip01(const float* hwc, int H, int W, float* chw, void* stream) {
    T3D_REQUIRE(hwc && chw, "NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)H * W;
    T3D_LAUNCH("hwc_to_chw_clip_kernel", st, hwc_to_chw_clip_kernel<<<grid_for(n), 256, 0, st>>>(hwc, n, chw));
    return T3D_OK;
}

}  // extern "C"
