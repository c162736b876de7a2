// t3d_sobel.cu -- Sobel thermal enhancer (sm_100a).
//
// Replaces /root/reference/thermal_dustr_model.py:110-142
// (ThermalDUSt3R.preprocess_thermal): per-(sample, channel) min/max normalise,
// depthwise 3x3 Sobel-x / Sobel-y with zero padding (conv2d padding=1, groups=3),
// mag = sqrt(ex^2 + ey^2), out = clamp((x_n + edge_weight * mag) * temp_scale, 0, 1).
// Backward for the two learnable scalars (edge_weight, temp_scale; :104-107).
#include "t3d_common.cuh"

namespace {

constexpr int kTS = 32;           // tile side
constexpr int kSobThreads = 256;

// ---- pass 1: per-plane min / max (deterministic: min/max are order independent)
__global__ void __launch_bounds__(256) plane_minmax_kernel(const float* __restrict__ x, int n, float* __restrict__ mm) {
    __shared__ float smin[8], smax[8];
    const int plane = blockIdx.x;
    const float* p = x + (size_t)plane * n;
    float lo = INFINITY, hi = -INFINITY;
    bool nan = false;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float v = __ldg(p + i);
        nan |= isnan(v);
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    if (nan) { lo = hi = __int_as_float(0x7fc00000); }       // torch.amin/amax propagate NaN
    // NaN-propagating warp reduction
    for (int o = 16; o > 0; o >>= 1) {
        const float a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = (isnan(a) || isnan(lo)) ? __int_as_float(0x7fc00000) : fminf(lo, a);
        hi = (isnan(b) || isnan(hi)) ? __int_as_float(0x7fc00000) : fmaxf(hi, b);
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            lo = (isnan(smin[w]) || isnan(lo)) ? __int_as_float(0x7fc00000) : fminf(lo, smin[w]);
            hi = (isnan(smax[w]) || isnan(hi)) ? __int_as_float(0x7fc00000) : fmaxf(hi, smax[w]);
        }
        mm[2 * plane] = lo; mm[2 * plane + 1] = hi;
    }
}

// ---- pass 2: normalise + Sobel + combine.  MODE 0: forward (writes out); MODE 1: backward for the two
// scalars (accumulates sum dout * d out/d edge_weight and d out/d temp_scale into per-block partials).
// in_channels == 1 with out planes 3: the single plane is computed once and written three times.
template <int MODE>
__global__ void __launch_bounds__(kSobThreads) sobel_enhance_kernel(const float* __restrict__ x, const float* __restrict__ mm,
                                                                    const float* __restrict__ params /* ew, ts */,
                                                                    int in_ch, int H, int W, int local_norm,
                                                                    float* __restrict__ out, const float* __restrict__ dout,
                                                                    float* __restrict__ partials) {
    __shared__ float t[kTS + 2][kTS + 2];
    __shared__ float red[kSobThreads / 32][2];
    const int tiles_x = (W + kTS - 1) / kTS, tiles_y = (H + kTS - 1) / kTS;
    const int plane = blockIdx.x / (tiles_x * tiles_y);          // b * in_ch + c
    const int tile = blockIdx.x - plane * tiles_x * tiles_y;
    const int i0 = (tile / tiles_x) * kTS, j0 = (tile % tiles_x) * kTS;
    const int b = plane / in_ch, c = plane - b * in_ch;
    const size_t n = (size_t)H * W;
    const float* p = x + (size_t)plane * n;
    const float ew = params[0], ts = params[1];
    float mn = 0.f, inv_r = 1.f, r = 1.f;
    if (local_norm) {
        mn = mm[2 * plane];
        r = (mm[2 * plane + 1] - mn) + 1e-6f;                     // :123-124
    }
    (void)inv_r;
    for (int q = threadIdx.x; q < (kTS + 2) * (kTS + 2); q += kSobThreads) {
        const int rr = q / (kTS + 2), cc = q - rr * (kTS + 2);
        const int i = i0 + rr - 1, j = j0 + cc - 1;
        float v = 0.f;                                            // zero padding of the NORMALISED image
        if (i >= 0 && i < H && j >= 0 && j < W) {
            v = __ldg(p + (size_t)i * W + j);
            if (local_norm) v = (v - mn) / r;
        }
        t[rr][cc] = v;
    }
    __syncthreads();
    const int out_rep = (in_ch == 1) ? 3 : 1;
    float s_ew = 0.f, s_ts = 0.f;
    for (int q = threadIdx.x; q < kTS * kTS; q += kSobThreads) {
        const int rr = q / kTS + 1, cc = q % kTS + 1;
        const int i = i0 + rr - 1, j = j0 + cc - 1;
        if (i >= H || j >= W) continue;
        // cross-correlation with [[-1,0,1],[-2,0,2],[-1,0,1]] and [[-1,-2,-1],[0,0,0],[1,2,1]] (:95-96)
        const float gx = (t[rr - 1][cc + 1] - t[rr - 1][cc - 1]) + 2.f * (t[rr][cc + 1] - t[rr][cc - 1]) +
                         (t[rr + 1][cc + 1] - t[rr + 1][cc - 1]);
        const float gy = (t[rr + 1][cc - 1] - t[rr - 1][cc - 1]) + 2.f * (t[rr + 1][cc] - t[rr - 1][cc]) +
                         (t[rr + 1][cc + 1] - t[rr - 1][cc + 1]);
        const float mag = sqrtf(gx * gx + gy * gy);               // :133
        const float y = t[rr][cc];
        const float pre = y + ew * mag;                           // :136
        const float e = pre * ts;                                 // :139
        if (MODE == 0) {
            const float o = isnan(e) ? e : fminf(fmaxf(e, 0.f), 1.f);   // :140
            for (int k = 0; k < out_rep; ++k)
                out[((size_t)(b * (in_ch * out_rep) + c * out_rep + k)) * n + (size_t)i * W + j] = o;
        } else {
            const bool pass = (e >= 0.f) && (e <= 1.f);           // clamp gradient mask (inclusive)
            float g = 0.f;
            for (int k = 0; k < out_rep; ++k)
                g += dout[((size_t)(b * (in_ch * out_rep) + c * out_rep + k)) * n + (size_t)i * W + j];
            if (pass) { s_ew += g * ts * mag; s_ts += g * pre; }
        }
    }
    if (MODE == 1) {
        s_ew = warp_sum(s_ew); s_ts = warp_sum(s_ts);
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s_ew; red[threadIdx.x >> 5][1] = s_ts; }
        __syncthreads();
        if (threadIdx.x < 2) {
            float v = 0.f;
            for (int w = 0; w < kSobThreads / 32; ++w) v += red[w][threadIdx.x];
            partials[(size_t)blockIdx.x * 2 + threadIdx.x] = v;
        }
    }
}

__global__ void sobel_bwd_finalize_kernel(const float* __restrict__ partials, int nblocks, float* __restrict__ dparams) {
    __shared__ double red[256];
    const int k = blockIdx.x;     // 0: d edge_weight, 1: d temp_scale
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) s += (double)partials[(size_t)i * 2 + k];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) dparams[k] = (float)red[0];
}

}  // namespace

extern "C" {

size_t t3d_sobel_workspace_bytes(int B, int C, int H, int W) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return 0;
    const size_t blocks = (size_t)B * C * ((W + kTS - 1) / kTS) * ((H + kTS - 1) / kTS);
    return t3d_align_up((size_t)B * C * 2 * sizeof(float), 256) + t3d_align_up(blocks * 2 * sizeof(float), 256);
}

static int sobel_common(int mode, const float* x, const float* params, int B, int C, int H, int W, int local_norm,
                        float* out, const float* dout, float* dparams, void* workspace, size_t ws_bytes, void* stream) {
    T3D_REQUIRE(x && params && workspace, "NULL pointer");
    T3D_REQUIRE(B >= 1 && (C == 1 || C == 3) && H >= 1 && W >= 1, "x must be [B,1|3,H,W]");
    T3D_REQUIRE((double)B * C * H * W < 2.0e9, "tensor too large");
    if (ws_bytes < t3d_sobel_workspace_bytes(B, C, H, W)) { t3d_set_error("workspace too small"); return T3D_ERR_WORKSPACE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* mm = reinterpret_cast<float*>(workspace);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + t3d_align_up((size_t)B * C * 2 * sizeof(float), 256));
    if (local_norm)
        T3D_LAUNCH("plane_minmax_kernel", st, plane_minmax_kernel<<<B * C, 256, 0, st>>>(x, H * W, mm));
    const int blocks = B * C * ((W + kTS - 1) / kTS) * ((H + kTS - 1) / kTS);
    if (mode == 0) {
        T3D_LAUNCH("sobel_enhance_kernel", st, sobel_enhance_kernel<0><<<blocks, kSobThreads, 0, st>>>(
            x, mm, params, C, H, W, local_norm, out, nullptr, nullptr));
    } else {
        T3D_LAUNCH("sobel_enhance_kernel", st, sobel_enhance_kernel<1><<<blocks, kSobThreads, 0, st>>>(
            x, mm, params, C, H, W, local_norm, nullptr, dout, partials));
        T3D_LAUNCH("sobel_bwd_finalize_kernel", st, sobel_bwd_finalize_kernel<<<2, 256, 0, st>>>(partials, blocks, dparams));
    }
    return T3D_OK;
}

int t3d_sobel_enhance_fwd(const float* x, const float* params, int B, int C, int H, int W, int local_norm,
                          float* out, void* workspace, size_t workspace_bytes, void* stream) {
    T3D_REQUIRE(out, "NULL pointer");
    return sobel_common(0, x, params, B, C, H, W, local_norm, out, nullptr, nullptr, workspace, workspace_bytes, stream);
}

int t3d_sobel_enhance_bwd_params(const float* x, const float* params, const float* dout, int B, int C, int H, int W,
                                 int local_norm, float* dparams, void* workspace, size_t workspace_bytes, void* stream) {
    T3D_REQUIRE(dout && dparams, "NULL pointer");
    return sobel_common(1, x, params, B, C, H, W, local_norm, nullptr, dout, dparams, workspace, workspace_bytes, stream);
}

}  // extern "C"
