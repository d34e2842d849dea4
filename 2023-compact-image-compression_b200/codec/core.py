"""Host-side mirror of the reference's codec entry points, backed by libcompact_b200.so.

Drop-in for `codec.core` of the reference (reference src/codec/core.py): the same two classes with the
same constructor signatures, return types and exceptions,

    Encoder(config, image, out_path=None).encode() -> bytes          (core.py:170-365)
    Decoder(config, file_bytes, out_path=None).decode() -> bytes     (core.py:367-543)

plus batch entry points on torch tensors / numpy arrays (`encode_batch`, `decode_batch`) that the
reference does not have.  All codec arithmetic runs in CUDA kernels behind the C ABI declared in
include/compact_b200.h; this module only validates arguments, moves buffers and raises the
reference's exceptions.  There is no CPU implementation here and none is called.
"""
from __future__ import annotations

import ctypes as C
import json
from collections import defaultdict

import numpy as np

from .. import _native as N

DEFAULT_CONFIG = {  # reference src/config.json
    "magic": "pact", "extension": "cct", "verbose": False, "block_size": 16,
    "encoder": {"channels": 1, "bytes_per_channel": 2, "end_of_file": 59, "deflate_compression": True,
                "transforms": {"delta": True, "zipper": False, "fractal": True, "segmentation": True}},
    "decoder": {"decode_format": "png"},
}


def default_config() -> dict:
    return json.loads(json.dumps(DEFAULT_CONFIG))


def _flags_from_config(config) -> int:
    enc = config["encoder"]
    if not enc["transforms"]["delta"]:
        raise NotImplementedError("Non-delta encoding not supported")  # core.py:221-222
    if enc["transforms"]["zipper"]:
        raise NotImplementedError("Zipper transform not supported or encouraged")  # core.py:224-225
    if config.get("block_size", 16) != 16:
        raise NotImplementedError("libcompact_b200 implements block_size 16 (config.json:9) only")
    if enc.get("channels", 1) != 1 or enc.get("bytes_per_channel", 2) != 2 or enc.get("end_of_file", 59) != 59 \
            or config.get("magic", "pact") != "pact":
        raise NotImplementedError("libcompact_b200 implements the shipped format constants only "
                                  "(magic 'pact', 1 channel, 2 bytes/channel, EOF 59)")
    return ((N.FLAG_FRACTAL if enc["transforms"]["fractal"] else 0)
            | (N.FLAG_SEGMENTATION if enc["transforms"]["segmentation"] else 0)
            | (N.FLAG_DEFLATE if enc["deflate_compression"] else 0))


def _torch():
    import torch
    return torch


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


# ---------------------------------------------------------------------------------------------- batch API

class _DeviceScratch:
    """Cached workspace tensors for the device-pointer entry points, one per (device, stream): calls on different
    streams may run at the same time and must not share scratch.  Results are never handed out from here."""
    cache: dict = {}

    @classmethod
    def get(cls, device, key: str, nbytes: int):
        torch = _torch()
        k = (str(device), int(torch.cuda.current_stream(device).cuda_stream), key)
        t = cls.cache.get(k)
        if t is None or t.numel() < nbytes:
            cls.cache[k] = t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return t

    @classmethod
    def clear(cls):
        cls.cache.clear()


def _ws_budget(device, want: int, one: int) -> int:
    torch = _torch()
    free, _ = torch.cuda.mem_get_info(device)
    return max(one, min(want, int(free * 0.5)))


def encode_batch(images, config=None, return_stats: bool = False):
    """Encode a stack of equally-shaped images.

    images: (B, rows, cols) uint16/int16 — a CUDA torch.Tensor (kernels run on its device, on torch's current
    stream, no host copies), or a CPU tensor / numpy array (goes through the host-buffer C entry point).
    Returns (arena, offsets[, stats]): file i is arena[offsets[i]:offsets[i+1]], byte-identical to the
    reference's Encoder output for image i.  Types follow the input (CUDA tensors in -> CUDA tensors out).
    """
    config = config or DEFAULT_CONFIG
    flags = _flags_from_config(config)
    L = N.lib()
    if _is_tensor(images) and images.is_cuda:
        torch = _torch()
        if images.dim() != 3 or images.element_size() != 2:
            raise ValueError("images must be (B, rows, cols) with a 2-byte dtype")
        images = images.contiguous()
        B, rows, cols = (int(s) for s in images.shape)
        dev = images.device
        with torch.cuda.device(dev):
            cap = max(B, 1) * int(L.cct_file_bound(rows, cols)) if rows * cols % 16 == 0 else 16
            one = int(L.cct_encode_workspace_bytes(1, rows, cols, flags))
            want = int(L.cct_encode_workspace_bytes(B, rows, cols, flags))
            arena = torch.empty(max(cap, 256), dtype=torch.uint8, device=dev)  # the caller's: a fresh tensor per call
            ws = _DeviceScratch.get(dev, "enc_ws", _ws_budget(dev, want, one)) if one else arena
            offsets = torch.empty(B + 1, dtype=torch.int64, device=dev)
            status = torch.empty(max(B, 1), dtype=torch.int32, device=dev)
            stats = torch.empty((max(B, 1), 4), dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(L.cct_encode_batch(images.data_ptr(), B, rows, cols, flags, arena.data_ptr(), arena.numel(),
                                       offsets.data_ptr(), status.data_ptr(), stats.data_ptr(), ws.data_ptr(),
                                       ws.numel(), stream))
        out = (arena, offsets, status[:B])
        return out + (stats[:B],) if return_stats else out
    arr = images.numpy() if _is_tensor(images) else np.asarray(images)
    if arr.ndim != 3 or arr.dtype.itemsize != 2:
        raise ValueError("images must be (B, rows, cols) with a 2-byte dtype")
    arr = np.ascontiguousarray(arr)
    B, rows, cols = arr.shape
    if (rows * cols) % 16:
        raise ValueError(f"cannot reshape array of size {rows * cols} into shape ({rows * cols // 16},16)")
    cap = max(B, 1) * int(L.cct_file_bound(rows, cols))
    arena = np.empty(cap, dtype=np.uint8)
    offsets = np.zeros(B + 1, dtype=np.int64)
    status = np.zeros(max(B, 1), dtype=np.uint32)
    stats = np.zeros((max(B, 1), 4), dtype=np.int32)
    N.check(L.cct_encode_host(arr.ctypes.data, B, rows, cols, flags, arena.ctypes.data, cap, offsets.ctypes.data,
                              status.ctypes.data, stats.ctypes.data, -1))
    out = (arena[:offsets[B]], offsets, status[:B])
    return out + (stats[:B],) if return_stats else out


def decode_batch(arena, offsets, shape):
    """Decode B .cct files stored back to back.  shape = (rows, cols) of every image (the header of each file
    is checked against it).  CUDA tensors in -> CUDA uint16 tensor (B, rows, cols) out; numpy in -> numpy out.
    Returns (images, status)."""
    L = N.lib()
    rows, cols = int(shape[0]), int(shape[1])
    if _is_tensor(arena) and arena.is_cuda:
        torch = _torch()
        dev = arena.device
        B = int(offsets.numel()) - 1
        offsets = offsets.to(device=dev, dtype=torch.int64).contiguous()
        arena = arena.contiguous()
        with torch.cuda.device(dev):
            one = int(L.cct_decode_workspace_bytes(1, rows, cols))
            want = int(L.cct_decode_workspace_bytes(B, rows, cols))
            ws = _DeviceScratch.get(dev, "dec_ws", _ws_budget(dev, want, one))
            px = torch.empty((B, rows, cols), dtype=torch.uint16, device=dev)
            status = torch.empty(max(B, 1), dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(L.cct_decode_batch(arena.data_ptr(), offsets.data_ptr(), B, rows, cols, px.data_ptr(),
                                       status.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return px, status[:B]
    buf = np.ascontiguousarray(np.frombuffer(arena, dtype=np.uint8) if isinstance(arena, (bytes, bytearray)) else arena)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    B = len(offsets) - 1
    px = np.empty((B, rows, cols), dtype=np.uint16)
    status = np.zeros(max(B, 1), dtype=np.uint32)
    N.check(L.cct_decode_host(buf.ctypes.data, offsets.ctypes.data, B, rows, cols, px.ctypes.data, status.ctypes.data, -1))
    return px, status[:B]


def read_header(file_bytes: bytes) -> dict:
    """Decoder.read_header (core.py:385-402) without touching the GPU."""
    r, c, ch, bpc, fl = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_uint32()
    rc = N.lib().cct_read_header(bytes(file_bytes[:13]), min(len(file_bytes), 13), C.byref(r), C.byref(c), C.byref(ch),
                                 C.byref(bpc), C.byref(fl))
    if rc != 0:
        raise ValueError("Image does not contain valid header")  # core.py:388-389
    return {"width": r.value, "height": c.value, "channels": ch.value, "bytes_per_channel": bpc.value,
            "fractal": bool(fl.value & 1), "segmentation": bool(fl.value & 2), "deflate": bool(fl.value & 4)}


def _raise_decode_status(st: int):
    if st & N.ST_BAD_MAGIC:
        raise ValueError("Image does not contain valid header")  # core.py:388-389
    if st & N.ST_INFLATE:
        import zlib  # exception type only: the reference lets zlib.error propagate (core.py:421)
        raise zlib.error("Error -3 while decompressing data: invalid stream")
    if st & N.ST_OVERFLOW:
        raise OverflowError("int too big to convert")  # core.py:506,516 recovered.to_bytes(2, ...)
    if st & (N.ST_TRUNCATED | N.ST_BAD_SHAPE):
        raise IndexError("index out of range")  # the reference runs off the end of its reader


# ---------------------------------------------------------------------------------------------- reference classes

class Encoder:
    """Same contract as the reference Encoder (core.py:170-365)."""

    def __init__(self, config, image, out_path=None):
        self.config = config
        self.image = image
        if _is_tensor(image):
            self.width, self.height = (int(s) for s in image.shape)
        else:
            self.image = np.asarray(image)
            self.width, self.height = self.image.shape  # width = rows (core.py:179)
        self.size = self.width * self.height
        self.out_path = out_path
        self.stats = [["Section", "Size (KB)", "Ratio (x)"]]
        self.info = defaultdict(int)
        self.header = b""
        self.domain_ok = True

    def encode(self) -> bytes:
        enc = self.config["encoder"]
        self.raw_size = self.size * enc["channels"] * enc["bytes_per_channel"]
        if self.raw_size > 400_000_000_000:
            raise MemoryError(f"Maximum byte count exceeded: {self.raw_size}")  # core.py:218-219
        _flags_from_config(self.config)  # NotImplementedError for !delta / zipper (core.py:221-225)
        if self.size % self.config["block_size"]:
            bs = self.config["block_size"]
            raise ValueError(f"cannot reshape array of size {self.size} into shape ({self.size // bs},{bs})")
        self.stats.append(["Original", self.raw_size / 1000, 1.0])
        img = self.image
        if _is_tensor(img):
            arena, offsets, status, stats = encode_batch(img.reshape(1, self.width, self.height), self.config, True)
            if img.is_cuda:
                n = int(offsets[1].item())
                output = bytes(arena[:n].cpu().numpy().tobytes())
                status, stats = status.cpu().numpy(), stats.cpu().numpy()
            else:
                output = arena.tobytes()
        else:
            if img.dtype.itemsize != 2:
                raise ValueError("libcompact_b200 encodes 2-byte pixels (config bytes_per_channel = 2)")
            arena, offsets, status, stats = encode_batch(img.reshape(1, self.width, self.height), self.config, True)
            output = arena.tobytes()
        self.info["delta"], self.info["full"] = int(stats[0][0]), int(stats[0][1])
        self.n_jumps, self.n_difficult = int(stats[0][2]), int(stats[0][3])
        self.domain_ok = not (int(status[0]) & N.ST_DOMAIN)
        self.header = output[:13]
        payload_len = self.info["delta"] + 2 * self.info["full"] + self.n_jumps + 1
        self.stats.append(["QOI", (13 + payload_len) / 1000, self.raw_size / (13 + payload_len)])
        if enc["deflate_compression"]:
            self.stats.append(["DEFLATE", len(output) / 1000, payload_len / (len(output) - 13)])
        self.stats.append(["Final", len(output) / 1000, self.raw_size / len(output)])
        if self.config.get("verbose"):
            from tabulate import tabulate
            print("\n" + json.dumps(self.info))
            print(tabulate(self.stats, headers="firstrow", tablefmt="simple_outline"))
        if self.out_path is not None:  # core.py:361-363
            with open(self.out_path, "wb") as fout:
                fout.write(output)
        return output


class Decoder:
    """Same contract as the reference Decoder (core.py:367-543)."""

    def __init__(self, config, file_bytes, out_path=None):
        self.config = config
        self.file_bytes = bytes(file_bytes)
        self.out_path = out_path
        self.fulls = []  # not reconstructed (debug aid in the reference, core.py:378,508)

    def read_header(self):
        h = read_header(self.file_bytes)
        self.width, self.height = h["width"], h["height"]
        self.channels, self.bytes_per_channel = h["channels"], h["bytes_per_channel"]
        self.fractal_transform, self.segmentation_transform = h["fractal"], h["segmentation"]
        self.deflate_compression = h["deflate"]

    def decode(self):
        self.read_header()
        self.size = self.width * self.height
        self.total_size = self.size * self.channels * self.bytes_per_channel
        if self.size % self.config["block_size"]:
            bs = self.config["block_size"]
            raise ValueError(f"cannot reshape array of size {self.size} into shape ({self.size // bs},{bs})")
        offsets = np.array([0, len(self.file_bytes)], dtype=np.int64)
        px, status = decode_batch(np.frombuffer(self.file_bytes, dtype=np.uint8), offsets, (self.width, self.height))
        _raise_decode_status(int(status[0]))
        output = px[0]
        if self.out_path is not None:  # core.py:522-540: 16-bit PNG preview of value << 4, returns the pixels
            preview = (output.astype(np.uint32) << 4).astype(np.uint16)
            from PIL import Image
            Image.fromarray(preview).save(self.out_path)
            return output
        return output.tobytes()  # core.py:543
