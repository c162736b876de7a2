/* t3d.h -- C ABI of libt3d_sm100.so, the B200 (sm_100a) implementation of the
 * Thermal3D-Vision per-pixel hot path.
 *
 * The reference has no FFI: its boundary is a set of Python module-level
 * functions.  Every entry point below names the
 * reference function (file:line under /root/reference) whose arithmetic it
 * replaces; the Python host side in thermal3d_vision_b200/ keeps the
 * reference's signatures and binds these symbols with ctypes.
 *
 * Conventions (all entry points):
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *    its name ends in _host;
 *  - the caller allocates every buffer (inputs, outputs, workspace); the
 *    library never allocates, frees or retains pointers;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *    implicit synchronisation, no host callbacks;
 *  - returns 0 (T3D_OK) on success, a negative t3d_status otherwise;
 *    t3d_last_error() gives a thread-local message;
 *  - tensors are dense, row-major, float32 unless stated.  Pointmaps are AoS
 *    [B,H,W,3]; thermal images planar [B,C,H,W]; confidences / depths [B,H,W].
 *    The 128-bit vector path is used when W % 4 == 0 and every pointer is
 *    16-byte aligned; otherwise the same kernels run their scalar-load
 *    variant (never a CPU fallback).
 */
#ifndef T3D_H_
#define T3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    T3D_OK = 0,
    T3D_ERR_BAD_ARG = -1,     /* bad shape / null pointer / unsupported option */
    T3D_ERR_WORKSPACE = -2,   /* workspace too small */
    T3D_ERR_CUDA = -3,        /* a CUDA runtime call or kernel launch failed */
    T3D_ERR_DEVICE = -4       /* not an sm_100 device */
} t3d_status;

#define T3D_ABI_VERSION 1

/* ------------------------------------------------------------------ library */
int t3d_version(void);
const char* t3d_last_error(void);
/* SM count and compute capability of the current device. */
int t3d_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Number of kernels this library has launched from the calling process
 * (monotonic; bench.py reports the delta over the timed region). */
uint64_t t3d_launch_count(void);

/* -------------------------------------------------------------------- loss */
/* Replaces utils/loss.py:75-98 (confidence_weighted_regression_loss) and
 * utils/loss.py:100-305 (enhanced_thermal_aware_loss) forward AND backward,
 * batched over B independent samples (the reference is called once per sample,
 * train_thermal_dustr.py:182-293; B = 1 reproduces that call).
 *
 * Layout of `out_sample` [B][8] float32, written by t3d_loss_fwd_bwd:
 *   0 total  1 basic  2 edge  3 smoothness  4 detail   (unweighted components,
 *   utils/loss.py:298-303)   5 valid (1.0 iff total is finite and > 0,
 *   train_thermal_dustr.py:320)   6,7 reserved.
 * Layout of `out_batch` [8] float32:
 *   0 mean total over valid samples (train_thermal_dustr.py:359)  1..4 mean
 *   components over valid samples  5 n_valid  6 B  7 reserved.
 * `out_sample_f64` (nullable) [B][8] double: the same per-sample numbers before
 * rounding to float32.
 */
#define T3D_LOSS_OUT_STRIDE 8

size_t t3d_loss_workspace_bytes(int B, int H, int W, int multi_scale);

/* Thermal-gradient statistics (utils/loss.py:184-201,239-249): for every
 * sample, view and scale, mean|Dx gray| and mean|Dy gray| over all h*w entries
 * of the zero-padded forward differences.  out_stats [B][2 views][2 scales][2]
 * float32 = the means (scale-2 slots are 0 when multi_scale == 0).
 * thermal_channels is 1 or 3 (gray = 0.299 c0 + 0.587 c1 + 0.114 c2 in fp32,
 * utils/loss.py:119-124). */
int t3d_thermal_grad_stats(const float* thermal1, const float* thermal2, int thermal_channels,
                           int B, int H, int W, int multi_scale,
                           float* out_stats, void* workspace, size_t workspace_bytes,
                           void* stream);

/* Fused forward + backward.  Launches: thermal statistics -> fused tile kernel
 * (loss partials + all gradients in one pass) -> deterministic second-stage
 * reduction.  conf1/conf2 may be NULL (-> ones, utils/loss.py:85-88);
 * thermal1/thermal2 may both be NULL (-> basic loss only, utils/loss.py:116);
 * dconf1/dconf2 may be NULL (confidence without grad).
 * Gradients are those of  grad_scale * loss_b  for every sample b (pass 1/B for
 * a batch mean); t3d_loss_rescale_invalid turns them into the gradients of the
 * mean over VALID samples without a host sync. */
int t3d_loss_fwd_bwd(const float* pred1, const float* pred2,
                     const float* gt1, const float* gt2,
                     const float* conf1, const float* conf2,
                     const float* thermal1, const float* thermal2, int thermal_channels,
                     float* dpred1, float* dpred2, float* dconf1, float* dconf2,
                     int B, int H, int W, int multi_scale,
                     float alpha, float edge_weight, float smoothness_weight, float detail_weight,
                     float grad_scale,
                     float* out_sample, float* out_batch, double* out_sample_f64,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Forward only (no gradient writes): same outputs as above. */
int t3d_loss_fwd(const float* pred1, const float* pred2,
                 const float* gt1, const float* gt2,
                 const float* conf1, const float* conf2,
                 const float* thermal1, const float* thermal2, int thermal_channels,
                 int B, int H, int W, int multi_scale,
                 float alpha, float edge_weight, float smoothness_weight, float detail_weight,
                 float* out_sample, float* out_batch, double* out_sample_f64,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Per-sample validity fix-up (train_thermal_dustr.py:320,359): multiplies the
 * gradients of sample b by  valid_b * B / n_valid  (read from out_sample /
 * out_batch on the device).  Every block exits immediately when all samples
 * are valid, so the common case costs one near-empty launch and no traffic. */
int t3d_loss_rescale_invalid(float* dpred1, float* dpred2, float* dconf1, float* dconf2,
                             const float* out_sample, const float* out_batch,
                             int B, int H, int W, void* stream);

/* Generic upstream gradient: grads *= *grad_output (a device scalar); no-op on
 * the device when *grad_output == 1.0f.  Used by autograd backward. */
int t3d_scale_grads(float* dpred1, float* dpred2, float* dconf1, float* dconf2,
                    const float* grad_output, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T3D_H_ */
