"""Compression-ratio study next to the codec: the reference's `scripts/evaluate.py` on arrays.

`comparison` of the reference (evaluate.py:53-95) records, per image, the size of the raw pixels, of
`zlib.compress(image.tobytes())` (default level, evaluate.py:69), of a 16-bit PNG (lib/png.py:4-15), of DICOM RLE
and JPEG 2000 (pydicom / an openjpeg executable: neither is in this image, so those two columns are left out) and
of the CCT file.  Here the CCT column comes from ONE batched GPU encode of all images (byte-exact with the
reference encoder, so the column is the reference's own); ZIP and PNG are host-side comparators exactly as the
reference computes them.  Rows and CSV layout follow evaluate.py:124-139.
"""
from __future__ import annotations

import io
import zlib

import numpy as np

FILE, RAW, ZIP, PNG, CCT = "File", "Raw", "ZIP", "PNG", "CCT"


def _png_size(image: np.ndarray) -> int:
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.ascontiguousarray(image).astype(np.uint16)).save(buf, format="PNG")  # mode I;16, as lib/png.py
    return buf.tell()


def comparison_batch(images, names=None, config=None, png: bool = True):
    """One row per image: {File, Raw, ZIP, [PNG,] CCT} in bytes.  `images`: uint16 [B, H, W] (numpy or torch)."""
    import torch
    from .codec.core import encode_batch
    t = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images).view(np.int16))
    _, offsets, _ = encode_batch(t.cuda() if not t.is_cuda else t, config=config)
    sizes = np.diff(offsets.cpu().numpy())
    host = t.cpu().numpy().view(np.uint16)
    rows = []
    for i in range(host.shape[0]):
        row = {FILE: names[i] if names else f"({i:04})", RAW: int(host[i].nbytes), ZIP: len(zlib.compress(host[i].tobytes()))}
        if png:
            row[PNG] = _png_size(host[i])
        row[CCT] = int(sizes[i])
        rows.append(row)
    return rows


def write_csv(rows, path: str) -> None:
    """The reference's results file: header line, then one comma-separated line per image (evaluate.py:129-133)."""
    with open(path, "w") as f:
        f.write(",".join(rows[0].keys()))
        for r in rows:
            f.write("\n" + ",".join(map(str, r.values())))


def ratios(rows):
    """Mean compression ratio (Raw / size) per method, as the reference's results notebook reports it."""
    out = {}
    for k in rows[0]:
        if k not in (FILE, RAW):
            out[k] = float(np.mean([r[RAW] / r[k] for r in rows]))
    return out
