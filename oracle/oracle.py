"""ctypes front end of the CPU oracle (oracle/cct_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does (tests/test_boundary.py greps for it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcct_oracle.so")
_lib = None


class Stats(C.Structure):
    _fields_ = [("n_delta", C.c_int64), ("n_full", C.c_int64), ("n_jump", C.c_int64),
                ("n_difficult", C.c_int64), ("payload_len", C.c_int64)]


def build(force: bool = False) -> str:
    """Compile oracle/libcct_oracle.so with gcc (needs zlib.h + libz, both in the image)."""
    srcs = [os.path.join(_HERE, f) for f in ("cct_oracle.c", "deflate9_port.c", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libcct_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, u16p, i32p, i64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_int32, C.c_int64))
        L.cct_oracle_curve.argtypes = [C.c_int, C.c_int, i32p]
        L.cct_oracle_curve.restype = None
        L.cct_oracle_payload_cap.argtypes = [C.c_int, C.c_int]
        L.cct_oracle_payload_cap.restype = C.c_int64
        L.cct_oracle_file_cap.argtypes = [C.c_int, C.c_int]
        L.cct_oracle_file_cap.restype = C.c_int64
        L.cct_oracle_encode.argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int64,
                                        u8p, i32p, i32p, C.POINTER(Stats)]
        L.cct_oracle_encode.restype = C.c_int64
        L.cct_oracle_decode.argtypes = [u8p, C.c_int64, u16p, C.c_int64, i32p]
        L.cct_oracle_decode.restype = C.c_int
        L.cct_oracle_encode_batch.argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p,
                                              C.c_int64, i64p, C.c_int]
        L.cct_oracle_encode_batch.restype = None
        L.cct_oracle_decode_batch.argtypes = [u8p, i64p, C.c_int, u16p, C.c_int64, C.POINTER(C.c_int), C.c_int]
        L.cct_oracle_decode_batch.restype = None
        L.cct_oracle_zlib_version.restype = C.c_char_p
        L.d9_compress.argtypes = [u8p, C.c_int64, u8p, C.c_int64, i64p]
        L.d9_compress.restype = C.c_int64
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def curve(width: int, height: int) -> np.ndarray:
    """Traversal order (curve.py:61-138): order[k] = raster index of the k-th visited pixel."""
    out = np.empty(width * height, dtype=np.int32)
    lib().cct_oracle_curve(width, height, _p(out, C.c_int32))
    return out


def encode(image: np.ndarray, fractal: bool = True, seg: bool = True, deflate: bool = True, detail: bool = False):
    """Oracle for Encoder(config, image).encode() (core.py:212-365).  Returns bytes, or
    (bytes, dict of intermediates) with detail=True."""
    img = np.ascontiguousarray(image)
    if img.dtype != np.uint16:
        img = img.astype(np.uint16)
    w, h = img.shape
    n = w * h
    if n % 16:
        raise ValueError(f"cannot reshape array of size {n} into shape ({n // 16},16)")
    L = lib()
    cap = int(L.cct_oracle_file_cap(w, h))
    out = np.empty(cap, dtype=np.uint8)
    st = Stats()
    if detail:
        payload = np.empty(int(L.cct_oracle_payload_cap(w, h)), dtype=np.uint8)
        order = np.empty(n, dtype=np.int32)
        jumps = np.empty(n // 16, dtype=np.int32)
        r = L.cct_oracle_encode(_p(img, C.c_uint16), w, h, int(fractal), int(seg), int(deflate), _p(out, C.c_uint8),
                                cap, _p(payload, C.c_uint8), _p(order, C.c_int32), _p(jumps, C.c_int32), C.byref(st))
    else:
        r = L.cct_oracle_encode(_p(img, C.c_uint16), w, h, int(fractal), int(seg), int(deflate), _p(out, C.c_uint8),
                                cap, None, None, None, C.byref(st))
    if r < 0:
        raise RuntimeError(f"oracle encode failed: {r}")
    data = out[:r].tobytes()
    if not detail:
        return data
    return data, {"payload": payload[:st.payload_len].tobytes(), "order": order, "jumps": jumps,
                  "n_delta": st.n_delta, "n_full": st.n_full, "n_jump": st.n_jump, "n_difficult": st.n_difficult}


def decode(file_bytes: bytes) -> np.ndarray:
    """Oracle for Decoder(config, file_bytes).decode() (core.py:404-543).  Returns uint16 (width, height)."""
    buf = np.frombuffer(file_bytes, dtype=np.uint8)
    dims = np.zeros(2, dtype=np.int32)
    L = lib()
    r = L.cct_oracle_decode(_p(buf, C.c_uint8), len(buf), None, 0, _p(dims, C.c_int32))
    if r == -1:
        raise ValueError("Image does not contain valid header")
    w, h = int(dims[0]), int(dims[1])
    out = np.empty(w * h, dtype=np.uint16)
    r = L.cct_oracle_decode(_p(buf, C.c_uint8), len(buf), _p(out, C.c_uint16), w * h, _p(dims, C.c_int32))
    if r == -4:
        raise OverflowError("reconstructed value outside [0, 65535]")
    if r != 0:
        raise RuntimeError(f"oracle decode failed: {r}")
    return out.reshape(w, h)


def encode_batch(images: np.ndarray, nthreads: int, fractal=True, seg=True, deflate=True):
    """Encode (B, W, H) uint16 on `nthreads` host threads.  Returns (arena u8, offsets int64[B+1])."""
    imgs = np.ascontiguousarray(images, dtype=np.uint16)
    B, w, h = imgs.shape
    L = lib()
    stride = int(L.cct_oracle_file_cap(w, h))
    out = np.empty(B * stride, dtype=np.uint8)
    lens = np.zeros(B, dtype=np.int64)
    L.cct_oracle_encode_batch(_p(imgs, C.c_uint16), B, w, h, int(fractal), int(seg), int(deflate),
                              _p(out, C.c_uint8), stride, _p(lens, C.c_int64), nthreads)
    if (lens < 0).any():
        raise RuntimeError("oracle batch encode failed")
    offsets = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    arena = np.empty(int(offsets[-1]), dtype=np.uint8)
    for i in range(B):
        arena[offsets[i]:offsets[i + 1]] = out[i * stride:i * stride + lens[i]]
    return arena, offsets


def encode_batch_timed(images: np.ndarray, nthreads: int):
    """Like encode_batch but returns only (lens, strided buffer) — no Python-side compaction (for timing)."""
    imgs = np.ascontiguousarray(images, dtype=np.uint16)
    B, w, h = imgs.shape
    L = lib()
    stride = int(L.cct_oracle_file_cap(w, h))
    out = np.empty(B * stride, dtype=np.uint8)
    lens = np.zeros(B, dtype=np.int64)

    def run():
        L.cct_oracle_encode_batch(_p(imgs, C.c_uint16), B, w, h, 1, 1, 1, _p(out, C.c_uint8), stride,
                                  _p(lens, C.c_int64), nthreads)
        return lens
    return run


def decode_batch(arena: np.ndarray, offsets: np.ndarray, shape, nthreads: int) -> np.ndarray:
    B = len(offsets) - 1
    w, h = shape
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    out = np.empty((B, w, h), dtype=np.uint16)
    rcs = (C.c_int * B)()
    lib().cct_oracle_decode_batch(_p(arena, C.c_uint8), _p(offsets, C.c_int64), B, _p(out, C.c_uint16), w * h,
                                  rcs, nthreads)
    if any(rcs):
        raise RuntimeError(f"oracle batch decode failed: {[r for r in rcs if r][:4]}")
    return out


def deflate9_port(data: bytes):
    """The from-scratch restatement of zlib-1.3 level 9 (oracle/deflate9_port.c); used to pin the GPU
    DEFLATE stage's intermediate results.  Returns the zlib stream."""
    src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
    cap = len(data) + len(data) // 1000 * 6 + 1024
    out = np.empty(cap, dtype=np.uint8)
    nsym = np.zeros(1, dtype=np.int64)
    r = lib().d9_compress(_p(src, C.c_uint8), len(data), _p(out, C.c_uint8), cap, _p(nsym, C.c_int64))
    if r < 0:
        raise RuntimeError(f"deflate9_port failed: {r}")
    return out[:r].tobytes()
