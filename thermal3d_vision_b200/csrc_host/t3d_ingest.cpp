// t3d_ingest.cpp -- host-side ingest of the training set's on-disk formats (include/t3d_ingest.h):
// 16-bit grayscale PNG thermal frames and .npy pseudo-GT arrays, decoded by a small thread pool straight
// into caller-provided (pinned) batch buffers.  zlib does the inflate; the PNG container, the five
// scanline filters and the .npy header are handled here (PNG specification, ISO/IEC 15948; NumPy NEP 1).
#include "../../include/t3d_ingest.h"

#include <zlib.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <exception>
#include <thread>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

struct Ihdr { int w, h, depth, color, interlace; };

int parse_ihdr(const uint8_t* d, size_t n, Ihdr* o) {
    if (n < 8 + 25 || memcmp(d, kPngSig, 8) != 0) return fail(T3D_INGEST_CORRUPT, "not a PNG (bad signature)");
    if (be32(d + 8) != 13 || memcmp(d + 12, "IHDR", 4) != 0) return fail(T3D_INGEST_CORRUPT, "PNG: first chunk is not IHDR");
    const uint8_t* p = d + 16;
    o->w = (int)be32(p); o->h = (int)be32(p + 4); o->depth = p[8]; o->color = p[9]; o->interlace = p[12];
    if (o->w <= 0 || o->h <= 0 || p[10] != 0 || p[11] != 0) return fail(T3D_INGEST_CORRUPT, "PNG: bad IHDR");
    return T3D_INGEST_OK;
}

inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Undo the scanline filter of `cur` (row bytes, in place) given the reconstructed previous row (or zeros).
int unfilter_row(int type, uint8_t* cur, const uint8_t* prev, int rowbytes, int bpp) {
    switch (type) {
        case 0: break;
        case 1: for (int i = bpp; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
        case 2: for (int i = 0; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + prev[i]); break;
        case 3:
            for (int i = 0; i < bpp; ++i) cur[i] = (uint8_t)(cur[i] + (prev[i] >> 1));
            for (int i = bpp; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + ((cur[i - bpp] + prev[i]) >> 1));
            break;
        case 4:
            for (int i = 0; i < bpp; ++i) cur[i] = (uint8_t)(cur[i] + prev[i]);
            for (int i = bpp; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + paeth(cur[i - bpp], prev[i], prev[i - bpp]));
            break;
        default: return fail(T3D_INGEST_CORRUPT, "PNG: unknown filter type %d", type);
    }
    return T3D_INGEST_OK;
}

int decode_png(const uint8_t* d, size_t n, uint16_t* out, int width, int height) {
    Ihdr ih = {0, 0, 0, 0, 0};
    if (int rc = parse_ihdr(d, n, &ih)) return rc;
    if (ih.color != 0 || (ih.depth != 16 && ih.depth != 8))
        return fail(T3D_INGEST_UNSUPPORTED, "PNG: colour type %d / bit depth %d (only grayscale 8/16-bit)", ih.color, ih.depth);
    if (ih.interlace != 0) return fail(T3D_INGEST_UNSUPPORTED, "PNG: interlaced images are not supported");
    if (ih.w != width || ih.h != height) return fail(T3D_INGEST_CORRUPT, "PNG is %dx%d, expected %dx%d", ih.w, ih.h, width, height);
    const int bpp = ih.depth / 8, rowbytes = ih.w * bpp;
    const size_t raw_size = (size_t)(rowbytes + 1) * ih.h;
    const size_t slack = 64;                       // a stream longer than the image shows up as produced > raw_size
    std::vector<uint8_t> raw(raw_size + slack);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) return fail(T3D_INGEST_CORRUPT, "zlib: inflateInit failed");
    zs.next_out = raw.data(); zs.avail_out = (uInt)(raw_size + slack);
    size_t pos = 8;
    bool done = false, saw_idat = false;
    int zrc = Z_OK;
    while (pos + 12 <= n && !done) {
        const uint32_t len = be32(d + pos);
        const uint8_t* type = d + pos + 4;
        if (pos + 12 + (size_t)len > n) { inflateEnd(&zs); return fail(T3D_INGEST_CORRUPT, "PNG: truncated chunk"); }
        if (memcmp(type, "IDAT", 4) == 0) {
            saw_idat = true;
            zs.next_in = const_cast<Bytef*>(d + pos + 8); zs.avail_in = len;
            while (zs.avail_in > 0 && zrc != Z_STREAM_END) {
                zrc = inflate(&zs, Z_NO_FLUSH);
                if (zrc != Z_OK && zrc != Z_STREAM_END) {
                    inflateEnd(&zs);
                    return fail(T3D_INGEST_CORRUPT, "PNG: inflate error %d (%s)", zrc, zs.msg ? zs.msg : "image data too long");
                }
            }
        } else if (memcmp(type, "IEND", 4) == 0) {
            done = true;
        }
        pos += 12 + (size_t)len;
    }
    const size_t produced = zs.total_out;
    inflateEnd(&zs);
    if (!saw_idat || produced != raw_size)
        return fail(T3D_INGEST_CORRUPT, "PNG: image data is %zu bytes, expected %zu", produced, raw_size);
    std::vector<uint8_t> zero((size_t)rowbytes, 0);
    const uint8_t* prev = zero.data();
    for (int y = 0; y < ih.h; ++y) {
        uint8_t* line = raw.data() + (size_t)y * (rowbytes + 1);
        if (int rc = unfilter_row(line[0], line + 1, prev, rowbytes, bpp)) return rc;
        prev = line + 1;
        uint16_t* o = out + (size_t)y * ih.w;
        if (bpp == 2) for (int x = 0; x < ih.w; ++x) o[x] = (uint16_t)((line[1 + 2 * x] << 8) | line[2 + 2 * x]);
        else for (int x = 0; x < ih.w; ++x) o[x] = line[1 + x];
    }
    return T3D_INGEST_OK;
}

int read_file(const char* path, std::vector<uint8_t>* buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail(T3D_INGEST_IO, "cannot open %s", path);
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return fail(T3D_INGEST_IO, "cannot stat %s", path); }
    buf->resize((size_t)sz);
    const size_t got = sz ? fread(buf->data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) return fail(T3D_INGEST_IO, "short read on %s", path);
    return T3D_INGEST_OK;
}

// ---------------------------------------------------------------- .npy
int parse_npy(const uint8_t* d, size_t n, char descr[16], int* fortran, int* ndim, int64_t shape[8], size_t* off) {
    if (n < 10 || memcmp(d, "\x93NUMPY", 6) != 0) return fail(T3D_INGEST_CORRUPT, "not a .npy file (bad magic)");
    const int major = d[6];
    size_t hlen, hoff;
    if (major == 1) { hlen = (size_t)d[8] | ((size_t)d[9] << 8); hoff = 10; }
    else if (major == 2 || major == 3) {
        if (n < 12) return fail(T3D_INGEST_CORRUPT, ".npy: truncated header");
        hlen = (size_t)d[8] | ((size_t)d[9] << 8) | ((size_t)d[10] << 16) | ((size_t)d[11] << 24); hoff = 12;
    } else return fail(T3D_INGEST_UNSUPPORTED, ".npy: format version %d", major);
    if (hoff + hlen > n) return fail(T3D_INGEST_CORRUPT, ".npy: truncated header");
    const std::string h(reinterpret_cast<const char*>(d + hoff), hlen);
    auto find_val = [&](const char* key) -> size_t {
        const size_t k = h.find(key);
        if (k == std::string::npos) return k;
        const size_t colon = h.find(':', k);
        return (colon == std::string::npos) ? colon : colon + 1;
    };
    size_t p = find_val("'descr'");
    if (p == std::string::npos) return fail(T3D_INGEST_CORRUPT, ".npy: no descr");
    const size_t q0 = h.find('\'', p);
    if (q0 == std::string::npos || h.find('[', p) < q0) return fail(T3D_INGEST_UNSUPPORTED, ".npy: structured dtype");
    const size_t q1 = h.find('\'', q0 + 1);
    if (q1 == std::string::npos || q1 - q0 - 1 >= 16) return fail(T3D_INGEST_CORRUPT, ".npy: bad descr");
    memcpy(descr, h.data() + q0 + 1, q1 - q0 - 1); descr[q1 - q0 - 1] = 0;
    p = find_val("'fortran_order'");
    if (p == std::string::npos) return fail(T3D_INGEST_CORRUPT, ".npy: no fortran_order");
    const size_t fo = h.find_first_not_of(' ', p);
    if (fo == std::string::npos) return fail(T3D_INGEST_CORRUPT, ".npy: bad fortran_order");
    *fortran = h.compare(fo, 4, "True") == 0 ? 1 : 0;
    p = find_val("'shape'");
    if (p == std::string::npos) return fail(T3D_INGEST_CORRUPT, ".npy: no shape");
    const size_t s0 = h.find('(', p), s1 = h.find(')', p);
    if (s0 == std::string::npos || s1 == std::string::npos || s1 < s0) return fail(T3D_INGEST_CORRUPT, ".npy: bad shape");
    *ndim = 0;
    size_t c = s0 + 1;
    while (c < s1) {
        while (c < s1 && (h[c] == ' ' || h[c] == ',')) ++c;
        if (c >= s1) break;
        if (*ndim >= 8) return fail(T3D_INGEST_UNSUPPORTED, ".npy: more than 8 dimensions");
        char* end = nullptr;
        const long long dim = strtoll(h.c_str() + c, &end, 10);
        if (end == h.c_str() + c) return fail(T3D_INGEST_CORRUPT, ".npy: bad shape");        // not a number: no progress
        if (dim < 0) return fail(T3D_INGEST_CORRUPT, ".npy: negative dimension");
        shape[(*ndim)++] = dim;
        c = (size_t)(end - h.c_str());
    }
    *off = hoff + hlen;
    return T3D_INGEST_OK;
}

inline float half_to_float(uint16_t hbits) {
    const uint32_t s = (uint32_t)(hbits & 0x8000u) << 16, e = (hbits >> 10) & 0x1fu, m = hbits & 0x3ffu;
    uint32_t bits;
    if (e == 0) {
        if (m == 0) bits = s;
        else {                                     // subnormal half -> normal float
            int sh = 0; uint32_t mm = m;
            while (!(mm & 0x400u)) { mm <<= 1; ++sh; }
            bits = s | ((uint32_t)(113 - sh) << 23) | ((mm & 0x3ffu) << 13);
        }
    } else if (e == 31) bits = s | 0x7f800000u | (m << 13);
    else bits = s | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &bits, 4);
    return f;
}

int read_npy_f32(const char* path, float* out, size_t elems) {
    std::vector<uint8_t> buf;
    if (int rc = read_file(path, &buf)) return rc;
    char descr[16]; int fortran, ndim; int64_t shape[8]; size_t off;
    if (int rc = parse_npy(buf.data(), buf.size(), descr, &fortran, &ndim, shape, &off)) return rc;
    size_t count = 1;
    for (int i = 0; i < ndim; ++i) {                        // overflow-checked product (a wrapped product could equal `elems`)
        const size_t dim = (size_t)shape[i];
        if (dim != 0 && count > SIZE_MAX / dim) return fail(T3D_INGEST_CORRUPT, "%s: shape overflows", path);
        count *= dim;
    }
    if (count != elems) return fail(T3D_INGEST_CORRUPT, "%s holds %zu elements, expected %zu", path, count, elems);
    if (fortran && ndim > 1) return fail(T3D_INGEST_UNSUPPORTED, "%s: Fortran-order arrays are not supported", path);
    const uint8_t* src = buf.data() + off;
    const size_t avail = buf.size() - off;
    auto is = [&](const char* a, const char* b) { return strcmp(descr, a) == 0 || strcmp(descr, b) == 0; };
    if (is("<f4", "=f4")) {
        if (avail < elems * 4) return fail(T3D_INGEST_CORRUPT, "%s: truncated data", path);
        memcpy(out, src, elems * 4);
    } else if (is("<f8", "=f8")) {
        if (avail < elems * 8) return fail(T3D_INGEST_CORRUPT, "%s: truncated data", path);
        for (size_t i = 0; i < elems; ++i) { double v; memcpy(&v, src + 8 * i, 8); out[i] = (float)v; }
    } else if (is("<f2", "=f2")) {
        if (avail < elems * 2) return fail(T3D_INGEST_CORRUPT, "%s: truncated data", path);
        for (size_t i = 0; i < elems; ++i) { uint16_t v; memcpy(&v, src + 2 * i, 2); out[i] = half_to_float(v); }
    } else return fail(T3D_INGEST_UNSUPPORTED, "%s: dtype %s (only <f4, <f8, <f2)", path, descr);
    return T3D_INGEST_OK;
}

template <typename Fn>
int run_pool(int count, int threads, int* status, Fn fn) {
    if (threads < 1) threads = 1;
    if (threads > count) threads = count;
    std::vector<int> st((size_t)count, 0);
    std::vector<std::string> msgs((size_t)count);
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= count) break;
            try {                                   // an exception escaping a worker thread would terminate the process
                st[i] = fn(i);
            } catch (const std::exception& e) {
                st[i] = fail(T3D_INGEST_CORRUPT, "item %d: %s", i, e.what());
            } catch (...) {
                st[i] = fail(T3D_INGEST_CORRUPT, "item %d: unknown exception", i);
            }
            if (st[i]) msgs[i] = g_err;            // thread-local message of the worker
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    int first = 0;
    for (int i = 0; i < count; ++i) {
        if (status) status[i] = st[i];
        if (st[i] && !first) { first = st[i]; snprintf(g_err, sizeof(g_err), "%s", msgs[i].c_str()); }
    }
    return first;
}

}  // namespace

extern "C" {

int t3d_ingest_version(void) { return 1; }
const char* t3d_ingest_last_error(void) { return g_err; }

int t3d_png_info(const uint8_t* data, size_t size, int* width, int* height, int* bit_depth, int* color_type, int* interlace) {
    if (!data) return fail(T3D_INGEST_BAD_ARG, "NULL pointer");
    Ihdr ih = {0, 0, 0, 0, 0};
    if (int rc = parse_ihdr(data, size, &ih)) return rc;
    if (width) *width = ih.w;
    if (height) *height = ih.h;
    if (bit_depth) *bit_depth = ih.depth;
    if (color_type) *color_type = ih.color;
    if (interlace) *interlace = ih.interlace;
    return T3D_INGEST_OK;
}

int t3d_png_decode_gray16(const uint8_t* data, size_t size, uint16_t* out, int width, int height) {
    if (!data || !out || width <= 0 || height <= 0) return fail(T3D_INGEST_BAD_ARG, "bad argument");
    return decode_png(data, size, out, width, height);
}

int t3d_png_decode_files_gray16(const char* const* paths, int count, uint16_t* out, int width, int height,
                                int threads, int* status) {
    if (!paths || !out || count < 0 || width <= 0 || height <= 0) return fail(T3D_INGEST_BAD_ARG, "bad argument");
    if (count == 0) return T3D_INGEST_OK;
    return run_pool(count, threads, status, [&](int i) -> int {
        std::vector<uint8_t> buf;
        if (int rc = read_file(paths[i], &buf)) return rc;
        return decode_png(buf.data(), buf.size(), out + (size_t)i * width * height, width, height);
    });
}

int t3d_npy_header(const uint8_t* data, size_t size, char descr[16], int* fortran_order, int* ndim, int64_t shape[8],
                   size_t* data_offset) {
    if (!data || !descr || !fortran_order || !ndim || !shape || !data_offset) return fail(T3D_INGEST_BAD_ARG, "NULL pointer");
    return parse_npy(data, size, descr, fortran_order, ndim, shape, data_offset);
}

int t3d_npy_read_files_f32(const char* const* paths, int count, float* out, size_t elems, int threads, int* status) {
    if (!paths || !out || count < 0) return fail(T3D_INGEST_BAD_ARG, "bad argument");
    if (count == 0) return T3D_INGEST_OK;
    return run_pool(count, threads, status, [&](int i) -> int { return read_npy_f32(paths[i], out + (size_t)i * elems, elems); });
}

}  // extern "C"
