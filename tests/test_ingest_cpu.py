"""CPU-only: libt3d_ingest.so (include/t3d_ingest.h) against the libraries the reference uses for the same files:
cv2.imread(path, cv2.IMREAD_ANYDEPTH) (data/dataset_loader.py:237-239) and np.load + .float() (:159-201).
PNGs are written both by OpenCV and by hand (every scanline filter type, split IDAT chunks)."""
import ctypes
import os
import re
import struct
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ingest():
    from thermal3d_vision_b200 import ingest as m
    m.build()
    return m


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png_gray(path, img, filters, idat_split=1, color_type=0, interlace=0):
    """Minimal PNG writer: img uint8/uint16 [H,W]; filters[y] in 0..4 chosen per row."""
    h, w = img.shape
    depth = 16 if img.dtype == np.uint16 else 8
    bpp = depth // 8
    rows = img.astype(">u2").tobytes() if depth == 16 else img.tobytes()
    rb = w * bpp
    raw = bytearray()
    prev = bytes(rb)
    for y in range(h):
        cur = rows[y * rb:(y + 1) * rb]
        f = filters[y % len(filters)]
        out = bytearray(rb)
        for i in range(rb):
            a = cur[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[f]
            out[i] = (cur[i] - pred) & 0xff
        raw.append(f)
        raw += out
        prev = cur
    comp = zlib.compress(bytes(raw), 6)
    parts = [comp[i * len(comp) // idat_split:(i + 1) * len(comp) // idat_split] for i in range(idat_split)]
    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n")
        fh.write(_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, interlace)))
        fh.write(_chunk(b"tEXt", b"Comment\x00thermal"))          # ancillary chunk: must be skipped
        for p in parts:
            fh.write(_chunk(b"IDAT", p))
        fh.write(_chunk(b"IEND", b""))


def test_header_symbols_are_exported_and_bound(ingest):
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "t3d_ingest.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(t3d_[a-z0-9_]+)\s*\(", text)))
    handle = ctypes.CDLL(ingest.LIB_PATH)
    assert not [s for s in syms if not hasattr(handle, s)]
    assert sorted(ingest.declared_symbols()) == syms


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8])
def test_every_filter_type_and_split_idat(ingest, tmp_path, dtype):
    rng = np.random.default_rng(1)
    hi = 65536 if dtype == np.uint16 else 256
    img = rng.integers(0, hi, (37, 53)).astype(dtype)
    img[5:9] = (np.arange(53) * 7 % hi).astype(dtype)                # smooth rows: non-trivial Sub / Paeth predictions
    paths = []
    for k, (filters, split) in enumerate([([0], 1), ([1], 1), ([2], 2), ([3], 1), ([4], 3), ([0, 1, 2, 3, 4], 5)]):
        p = str(tmp_path / f"f{k}.png")
        write_png_gray(p, img, filters, idat_split=split)
        paths.append(p)
    out = ingest.read_thermal_png_batch(paths, threads=3, pin=False).numpy()
    assert out.dtype == np.uint16 and out.shape == (6, 37, 53)
    for k in range(6):
        assert np.array_equal(out[k], img.astype(np.uint16)), k
    cv2 = pytest.importorskip("cv2")
    ref = cv2.imread(paths[5], cv2.IMREAD_ANYDEPTH)                 # what the reference's loader returns
    assert np.array_equal(out[5], ref.astype(np.uint16))


def test_matches_opencv_on_full_size_thermal_frames(ingest, tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    frames = [(22800 + 400 * rng.standard_normal((512, 640))).clip(0, 65535).astype(np.uint16) for _ in range(5)]
    frames.append(rng.integers(0, 65536, (512, 640)).astype(np.uint16))
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"t{i}.png")
        assert cv2.imwrite(p, f)
        paths.append(p)
    assert ingest.png_info(paths[0]) == (640, 512, 16, 0, 0)
    out = ingest.read_thermal_png_batch(paths, threads=4, pin=False)
    for i, p in enumerate(paths):
        assert np.array_equal(out[i].numpy(), cv2.imread(p, cv2.IMREAD_ANYDEPTH)), i
    again = ingest.read_thermal_png_batch(paths, threads=1, pin=False, out=out.clone())
    assert np.array_equal(again.numpy(), out.numpy())


def test_png_errors(ingest, tmp_path):
    img = np.arange(48, dtype=np.uint16).reshape(6, 8)
    good = str(tmp_path / "good.png"); write_png_gray(good, img, [4])
    other = str(tmp_path / "other.png"); write_png_gray(other, np.zeros((6, 9), np.uint16), [0])
    rgb = str(tmp_path / "rgb.png"); write_png_gray(rgb, img, [0], color_type=2)
    lace = str(tmp_path / "lace.png"); write_png_gray(lace, img, [0], interlace=1)
    trunc = str(tmp_path / "trunc.png"); open(trunc, "wb").write(open(good, "rb").read()[:60])
    junk = str(tmp_path / "junk.png"); open(junk, "wb").write(b"not a png at all, just bytes" * 4)
    for bad, word in ((other, "expected"), (rgb, "colour type"), (lace, "interlaced"), (trunc, ""), (junk, "signature"),
                      (str(tmp_path / "missing.png"), "open")):
        with pytest.raises(ingest.IngestError) as e:
            ingest.read_thermal_png_batch([good, bad], pin=False)
        assert word in str(e.value)
    with pytest.raises(ValueError):
        ingest.read_thermal_png_batch([], pin=False)


def test_npy_matches_numpy(ingest, tmp_path):
    rng = np.random.default_rng(3)
    shape = (24, 32, 3)
    arrays = [rng.standard_normal(shape).astype(np.float32), rng.standard_normal(shape),
              rng.standard_normal(shape).astype(np.float16)]
    paths = []
    for i, a in enumerate(arrays):
        p = str(tmp_path / f"a{i}.npy"); np.save(p, a); paths.append(p)
    p2 = str(tmp_path / "v2.npy")
    with open(p2, "wb") as fh:
        np.lib.format.write_array(fh, arrays[0], version=(2, 0))
    paths.append(p2)
    out = ingest.read_npy_batch_f32(paths, shape, threads=2, pin=False).numpy()
    import torch
    for i, a in enumerate(arrays + [arrays[0]]):
        assert np.array_equal(out[i], torch.from_numpy(np.load(paths[i])).float().numpy()), i
    descr, fortran, shp, off = ingest.npy_header(paths[1])
    assert (descr, fortran, shp) == ("<f8", False, shape)
    assert np.array_equal(np.memmap(paths[1], dtype="<f8", mode="r", offset=off, shape=shp), arrays[1])
    # errors: wrong element count, Fortran order, integer dtype
    with pytest.raises(ingest.IngestError):
        ingest.read_npy_batch_f32(paths[:1], (24, 32), pin=False)
    pf = str(tmp_path / "f.npy"); np.save(pf, np.asfortranarray(arrays[0]))
    with pytest.raises(ingest.IngestError):
        ingest.read_npy_batch_f32([pf], shape, pin=False)
    pi = str(tmp_path / "i.npy"); np.save(pi, np.zeros(shape, np.int32))
    with pytest.raises(ingest.IngestError):
        ingest.read_npy_batch_f32([pi], shape, pin=False)


def test_malformed_npy_headers_are_errors_not_crashes(ingest, tmp_path):
    """Hand-made corrupt .npy headers: every one comes back as an IngestError from the worker pool (no exception
    escapes a worker thread, no wrapped shape product passes for the expected element count)."""
    def write(name, header: bytes, payload: bytes = b"\\0" * 64):
        hdr = header + b" " * ((64 - (10 + len(header) + 1) % 64) % 64) + b"\\n"
        p = str(tmp_path / name)
        with open(p, "wb") as fh:
            fh.write(b"\\x93NUMPY\\x01\\x00" + len(hdr).to_bytes(2, "little") + hdr + payload)
        return p
    bad = [
        write("no_colon.npy", b"{'descr' '<f4', 'fortran_order': False, 'shape': (4,), }"),
        write("fortran_blank.npy", b"{'descr': '<f4', 'shape': (4,), 'fortran_order':"),
        write("neg.npy", b"{'descr': '<f4', 'fortran_order': False, 'shape': (-4, -1), }"),
        write("junk_shape.npy", b"{'descr': '<f4', 'fortran_order': False, 'shape': (x, y), }"),
        write("overflow.npy", b"{'descr': '<f4', 'fortran_order': False, 'shape': (4294967296, 4294967296, 4), }"),
        write("no_shape_close.npy", b"{'descr': '<f4', 'fortran_order': False, 'shape': (4, "),
    ]
    for p in bad:
        with pytest.raises(ingest.IngestError):
            ingest.read_npy_batch_f32([p, p], (4,), threads=2, pin=False)
