/* t3d_ingest.h -- C ABI of libt3d_ingest.so: host-side raw-frame ingest for the Thermal3D-Vision hot path
 * (SURVEY.md section 8f row 3).  Plain C, no CUDA: it fills caller-provided (ideally pinned) host buffers
 * that are then copied to the device in one cudaMemcpyAsync per batch and handed to t3d_preprocess_train_u16.
 *
 * Replaces, for the data formats the training set uses:
 *   - cv2.imread(path, cv2.IMREAD_ANYDEPTH) of the 16-bit grayscale thermal PNGs
 *     (data/dataset_loader.py:237-239, thermal_dustr_inference.py:30-33): non-interlaced PNG, colour type 0,
 *     bit depth 16 (big-endian samples -> host uint16) or 8 (widened to uint16 without scaling);
 *   - np.load(path) of the pseudo-GT arrays followed by .float() (data/dataset_loader.py:159-201): .npy format
 *     1.0 / 2.0 / 3.0, C order, little-endian float32 / float64 / float16 -> float32.
 * Everything else (interlaced or colour PNGs, Fortran-order or object arrays) returns T3D_INGEST_UNSUPPORTED so
 * the caller can report it; there is no silent conversion.
 *
 * All functions return 0 on success or a negative t3d_ingest_status; t3d_ingest_last_error() gives a
 * thread-local message.  Thread-safe; the *_files_* functions decode with `threads` worker threads. */
#ifndef T3D_INGEST_H_
#define T3D_INGEST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    T3D_INGEST_OK = 0,
    T3D_INGEST_BAD_ARG = -1,
    T3D_INGEST_IO = -2,            /* file missing / short read */
    T3D_INGEST_CORRUPT = -3,       /* not a PNG / .npy, bad chunk, inflate error, size mismatch */
    T3D_INGEST_UNSUPPORTED = -4    /* valid file, format outside the scope above */
} t3d_ingest_status;

int t3d_ingest_version(void);
const char* t3d_ingest_last_error(void);

/* IHDR of a PNG held in memory. */
int t3d_png_info(const uint8_t* data, size_t size, int* width, int* height, int* bit_depth, int* color_type,
                 int* interlace);
/* Decode a grayscale 8/16-bit PNG held in memory into out[height][width] (host-endian uint16). */
int t3d_png_decode_gray16(const uint8_t* data, size_t size, uint16_t* out, int width, int height);
/* Decode `count` PNG files of identical size into out[count][height][width] with `threads` workers.
 * status[i] (nullable) receives the per-file status; the return value is the first failure (or 0). */
int t3d_png_decode_files_gray16(const char* const* paths, int count, uint16_t* out, int width, int height,
                                int threads, int* status);

/* Header of a .npy file held in memory: descr (NUL-terminated, e.g. "<f4"), Fortran flag, shape, and the
 * byte offset of the data. */
int t3d_npy_header(const uint8_t* data, size_t size, char descr[16], int* fortran_order, int* ndim,
                   int64_t shape[8], size_t* data_offset);
/* Read `count` .npy files, each holding exactly `elems` elements of <f4 / <f8 / <f2 in C order, into
 * out[count][elems] as float32 (np.load(...).float()) with `threads` workers. */
int t3d_npy_read_files_f32(const char* const* paths, int count, float* out, size_t elems, int threads,
                           int* status);

#ifdef __cplusplus
}
#endif
#endif  /* T3D_INGEST_H_ */
