"""ctypes binding of libcompact_b200.so (include/compact_b200.h).

The shared library holds every kernel of the hot path.  There is NO CPU fallback: if the library is
missing or no CUDA device is present, calls raise instead of silently computing elsewhere.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcompact_b200.so")
CSRC = os.path.join(_HERE, "csrc")

FLAG_FRACTAL, FLAG_SEGMENTATION, FLAG_DEFLATE = 1, 2, 4
ST_DOMAIN, ST_CAPACITY, ST_BAD_MAGIC, ST_BAD_SHAPE, ST_INFLATE, ST_OVERFLOW, ST_TRUNCATED = 1, 2, 4, 8, 16, 32, 64
E_ARG, E_SHAPE, E_WORKSPACE, E_CUDA, E_TOO_LARGE, E_NO_DEVICE, E_CAPACITY = -1, -2, -3, -4, -5, -6, -7

_lib = None


class NativeError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"libcompact_b200: {detail}" if detail else f"libcompact_b200 error {code}")


def build(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "compact_b200.h")]
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        r = subprocess.run(["make", "-C", CSRC, "-B"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc build of libcompact_b200.so failed:\n" + r.stdout[-4000:])
    return LIB_PATH


def lib():
    """Load the library (building it first if the .so is absent).  Raises ImportError when it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            build()
        except Exception as e:  # no nvcc, compile error ...
            raise ImportError(f"libcompact_b200.so is missing and could not be built ({e}); "
                              "run `python -c 'import __graft_entry__ as g; g.build()'`") from e
    L = C.CDLL(LIB_PATH)
    vp, u8p, u16p = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses: device or host pointers
    sz = C.c_size_t
    L.cct_version.restype = C.c_int
    L.cct_strerror.restype = C.c_char_p
    L.cct_strerror.argtypes = [C.c_int]
    L.cct_last_cuda_error.restype = C.c_char_p
    L.cct_device_count.restype = C.c_int
    L.cct_file_bound.restype = sz
    L.cct_file_bound.argtypes = [C.c_int, C.c_int]
    L.cct_payload_bound.restype = sz
    L.cct_payload_bound.argtypes = [C.c_int, C.c_int]
    L.cct_encode_workspace_bytes.restype = sz
    L.cct_encode_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32]
    L.cct_decode_workspace_bytes.restype = sz
    L.cct_decode_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.cct_encode_batch.restype = C.c_int
    L.cct_encode_batch.argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_uint32, u8p, sz, vp, vp, vp, vp, sz, vp]
    L.cct_decode_batch.restype = C.c_int
    L.cct_decode_batch.argtypes = [u8p, vp, C.c_int, C.c_int, C.c_int, u16p, vp, vp, sz, vp]
    L.cct_read_header.restype = C.c_int
    L.cct_read_header.argtypes = [C.c_char_p, sz, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
    L.cct_encode_host.restype = C.c_int
    L.cct_encode_host.argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_uint32, u8p, sz, vp, vp, vp, C.c_int]
    L.cct_decode_host.restype = C.c_int
    L.cct_decode_host.argtypes = [u8p, vp, C.c_int, C.c_int, C.c_int, u16p, vp, C.c_int]
    L.cct_release.restype = None
    L.cct_debug_encode_stages.restype = C.c_int
    L.cct_debug_encode_stages.argtypes = [u16p, C.c_int, C.c_int, C.c_uint32, vp, vp, vp, vp, vp, sz, vp]
    L.cct_deflate_workspace_bytes.restype = sz
    L.cct_deflate_workspace_bytes.argtypes = [C.c_int, sz]
    L.cct_deflate_bound.restype = sz
    L.cct_deflate_bound.argtypes = [sz]
    L.cct_deflate_batch.restype = C.c_int
    L.cct_deflate_batch.argtypes = [u8p, sz, vp, C.c_int, sz, u8p, sz, vp, vp, sz, vp]
    L.cct_inflate_batch.restype = C.c_int
    L.cct_inflate_batch.argtypes = [u8p, sz, vp, C.c_int, u8p, sz, vp, vp, vp, sz, vp]
    L.cct_debug_host_schedule.restype = C.c_int
    L.cct_debug_host_schedule.argtypes = [C.c_int, sz, C.c_int, vp, C.c_int]
    L.cct_debug_match_tables.restype = C.c_int
    L.cct_debug_match_tables.argtypes = [u8p, sz, vp, C.c_int, sz, vp, vp, vp, sz, vp, sz, vp]
    L.cct_inflate_workspace_bytes.restype = sz
    L.cct_inflate_workspace_bytes.argtypes = [C.c_int, sz]
    L.cct_launch_count.restype = C.c_longlong
    L.cct_launch_count.argtypes = [C.c_int]
    L.cct_profile_enable.restype = C.c_int
    L.cct_profile_enable.argtypes = [C.c_int]
    L.cct_profile_select.restype = C.c_int
    L.cct_profile_select.argtypes = [C.c_char_p]
    L.cct_profile_read.restype = C.c_int
    L.cct_profile_read.argtypes = [C.c_char_p, sz]
    _lib = L
    return L


def profile_read() -> dict:
    """Per-kernel {name: (launches, total_ms)} recorded since cct_profile_enable(1); synchronises the device."""
    buf = C.create_string_buffer(1 << 16)
    check(lib().cct_profile_read(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split()
        out[name] = (int(n), float(ms))
    return out


EXPORTS = ["cct_version", "cct_strerror", "cct_last_cuda_error", "cct_device_count", "cct_file_bound",
           "cct_encode_workspace_bytes", "cct_decode_workspace_bytes", "cct_encode_batch", "cct_decode_batch",
           "cct_read_header", "cct_encode_host", "cct_decode_host", "cct_release", "cct_payload_bound",
           "cct_debug_encode_stages", "cct_deflate_workspace_bytes", "cct_deflate_bound", "cct_deflate_batch",
           "cct_inflate_batch", "cct_inflate_workspace_bytes", "cct_debug_match_tables", "cct_debug_host_schedule", "cct_launch_count", "cct_profile_enable", "cct_profile_select", "cct_profile_read"]


def check(code: int):
    """Map a C-ABI return code to the exception the reference raises for the same condition."""
    if code == 0:
        return
    L = lib()
    msg = L.cct_strerror(code).decode()
    if code == E_CUDA:
        msg += ": " + L.cct_last_cuda_error().decode()
    if code == E_TOO_LARGE:
        raise MemoryError(msg)  # core.py:218-219
    if code == E_SHAPE:
        raise ValueError(msg)  # core.py:245 (numpy reshape)
    raise NativeError(code, msg)
