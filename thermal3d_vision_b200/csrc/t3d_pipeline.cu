// t3d_pipeline.cu -- packs the per-step scalars of the hot path into the one
// small vector that is all-reduced across ranks:
// training [sum valid loss, sum basic, edge, smooth, detail, n_valid, B] as
// train_thermal_dustr.py:320,359 accumulates them, evaluation [7 metric sums,
// n_images] as utils/metrics.py:128-136 does (non-finite metrics skipped but
// the image still counted).
#include "t3d_common.cuh"

namespace {

__global__ void __launch_bounds__(128) pack_step_result_kernel(const float* __restrict__ loss_per_sample,
                                                               const double* __restrict__ metrics_f64, int B,
                                                               int n_images, double* __restrict__ out) {
    __shared__ double red[4][14];
    const int tid = threadIdx.x;
    double v[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) v[k] = 0.0;
    if (loss_per_sample) {
        for (int b = tid; b < B; b += 128) {
            const float* o = loss_per_sample + (size_t)b * T3D_LOSS_OUT_STRIDE;
            if (o[5] != 0.f) {
#pragma unroll
                for (int k = 0; k < 5; ++k) v[k] += (double)o[k];
                v[5] += 1.0;
            }
        }
    }
    if (metrics_f64) {
        for (int b = tid; b < n_images; b += 128) {
            const double* m = metrics_f64 + (size_t)b * 8;
#pragma unroll
            for (int k = 0; k < 7; ++k) if (isfinite(m[k])) v[7 + k] += m[k];
        }
    }
    // fixed butterfly per warp, fixed order over the 4 warps: deterministic
    const int lane = tid & 31, wrp = tid >> 5;
#pragma unroll
    for (int k = 0; k < 14; ++k) {
        const double r = warp_sum(v[k]);
        if (lane == 0) red[wrp][k] = r;
    }
    __syncthreads();
    if (tid < 14) red[0][tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    __syncthreads();
    if (tid < 16) {
        double r;
        if (tid == 6) r = loss_per_sample ? (double)B : 0.0;
        else if (tid == 14) r = metrics_f64 ? (double)n_images : 0.0;
        else if (tid == 15) r = 0.0;
        else r = red[0][tid];
        out[tid] = r;
    }
}

// pack + the validity fix-up of the loss gradients (scale_grads_kernel<0> of t3d_loss.cu) as ONE launch: block 0
// packs, every block then checks the batch's valid count and returns at once when all samples are valid.
struct EpilogueArgs {
    float* dpred[2]; float* dconf[2];
    const float* out_sample; const float* out_batch; const double* metrics_f64;
    int B, n_images; size_t plane; double* out16;
};

__global__ void __launch_bounds__(128) step_epilogue_kernel(const EpilogueArgs a) {
    if (blockIdx.x == 0) {
        __shared__ double red[4][14];
        const int tid = threadIdx.x;
        double v[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) v[k] = 0.0;
        for (int b = tid; b < a.B; b += 128) {
            const float* o = a.out_sample + (size_t)b * T3D_LOSS_OUT_STRIDE;
            if (o[5] != 0.f) {
#pragma unroll
                for (int k = 0; k < 5; ++k) v[k] += (double)o[k];
                v[5] += 1.0;
            }
        }
        if (a.metrics_f64) {
            for (int b = tid; b < a.n_images; b += 128) {
                const double* m = a.metrics_f64 + (size_t)b * 8;
#pragma unroll
                for (int k = 0; k < 7; ++k) if (isfinite(m[k])) v[7 + k] += m[k];
            }
        }
        const int lane = tid & 31, wrp = tid >> 5;
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            const double r = warp_sum(v[k]);
            if (lane == 0) red[wrp][k] = r;
        }
        __syncthreads();
        if (tid < 16) {
            double r;
            if (tid == 6) r = (double)a.B;
            else if (tid == 14) r = a.metrics_f64 ? (double)a.n_images : 0.0;
            else if (tid == 15) r = 0.0;
            else r = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
            a.out16[tid] = r;
        }
    }
    const float nv = a.out_batch[5];
    if (nv == (float)a.B) return;                      // every sample valid: the a-priori 1/B scale is right
    const size_t per_sample[2] = {a.plane * 3, a.plane};
#pragma unroll
    for (int which = 0; which < 4; ++which) {
        float* base = (which < 2) ? a.dpred[which] : a.dconf[which - 2];
        if (!base) continue;
        const size_t n = per_sample[which >> 1];
        const size_t total = n * a.B;
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
            const int b = (int)(idx / n);
            const bool valid = a.out_sample[(size_t)b * T3D_LOSS_OUT_STRIDE + 5] != 0.f;
            const float f = (valid && nv > 0.f) ? (float)a.B / nv : 0.f;
            base[idx] = (f == 0.f) ? 0.f : base[idx] * f;          // invalid samples may hold NaN / Inf: exact zeros
        }
    }
}

}  // namespace

extern "C" int t3d_step_epilogue(float* dpred1, float* dpred2, float* dconf1, float* dconf2,
                                 const float* loss_per_sample, const float* loss_batch, const double* metrics_f64,
                                 int B, int H, int W, int n_images, double* out16, void* stream) {
    T3D_REQUIRE(loss_per_sample && loss_batch && out16, "NULL pointer");
    T3D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && n_images >= 0, "bad dims");
    EpilogueArgs a;
    a.dpred[0] = dpred1; a.dpred[1] = dpred2; a.dconf[0] = dconf1; a.dconf[1] = dconf2;
    a.out_sample = loss_per_sample; a.out_batch = loss_batch; a.metrics_f64 = metrics_f64;
    a.B = B; a.n_images = n_images; a.plane = (size_t)H * W; a.out16 = out16;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    T3D_LAUNCH("step_epilogue_kernel", st, step_epilogue_kernel<<<t3d_sm_count() * 8, 128, 0, st>>>(a));
    return T3D_OK;
}

extern "C" int t3d_pack_step_result(const float* loss_per_sample, const double* metrics_f64, int B, int n_images,
                                    double* out16, void* stream) {
    T3D_REQUIRE(out16, "NULL pointer");
    T3D_REQUIRE(B >= 0 && n_images >= 0, "bad dims");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    T3D_LAUNCH("pack_step_result_kernel", st,
               pack_step_result_kernel<<<1, 128, 0, st>>>(loss_per_sample, metrics_f64, B, n_images, out16));
    return T3D_OK;
}
