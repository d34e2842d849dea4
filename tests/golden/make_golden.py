"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (all committed):
  reference_testing.cct   verbatim copy of the reference's own golden vector data/working/testing.cct
  golden_small.npz        inputs + reference Encoder outputs for small images (shapes, flag variants, quirks)
  golden_digests.json     sha256/length of reference Encoder outputs for synth_ct(seed) 512x512 slices and the
                          2048x2048 radiograph, the sha1 of the golden decode, and the check that
                          decoded-testing.png >> 4 equals it
The reference is imported from its mount; nothing of its source is copied.
"""
import hashlib
import importlib
import json
import os
import shutil
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, ROOT)
from codec.core import Decoder, Encoder  # noqa: E402  (the reference)

synth = importlib.import_module("2023-compact-image-compression_b200.synth")


def config(fractal=True, seg=True, deflate=True):
    cfg = json.load(open(os.path.join(REF, "src", "config.json")))
    cfg["verbose"] = False
    cfg["encoder"]["transforms"]["fractal"] = fractal
    cfg["encoder"]["transforms"]["segmentation"] = seg
    cfg["encoder"]["deflate_compression"] = deflate
    return cfg


def ref_encode(img, **kw):
    return Encoder(config(**kw), img, None).encode()


def ref_decode(data, shape):
    return np.frombuffer(Decoder(config(), data, None).decode(), dtype=np.uint16).reshape(shape)


def main():
    digests = {}
    # 1. the reference's own golden vector
    shutil.copyfile(os.path.join(REF, "data/working/testing.cct"), os.path.join(HERE, "reference_testing.cct"))
    g = open(os.path.join(HERE, "reference_testing.cct"), "rb").read()
    img = ref_decode(g, (512, 512))
    assert ref_encode(img) == g
    from PIL import Image
    png = np.asarray(Image.open(os.path.join(REF, "data/working/decoded-testing.png"))).astype(np.uint16)
    digests["reference_testing"] = {
        "file_sha256": hashlib.sha256(g).hexdigest(), "file_len": len(g),
        "decoded_sha1": hashlib.sha1(img.tobytes()).hexdigest(),
        "png_shr4_equals_decode": bool(((png >> 4) == img).all()),
    }

    # 2. small images through the reference encoder (and decoder where inside the lossless domain)
    small = {}
    cases = []
    rng = np.random.default_rng(1234)
    for k, (h, w) in enumerate([(4, 4), (16, 16), (12, 12), (20, 20), (32, 64), (64, 32), (48, 80), (40, 24),
                                (64, 64), (128, 128), (96, 160), (8, 2), (2, 8), (1, 16), (16, 1), (36, 44)]):
        cases.append((f"walk_{h}x{w}", synth.random_walk_image(100 + k, h, w), True))
    for k in range(4):  # out-of-domain uint16 noise: encoder parity only
        cases.append((f"noise16_{k}", rng.integers(0, 65536, size=(32, 32)).astype(np.uint16), False))
    b0 = synth.random_walk_image(7, 32, 32)
    b0[:4, :4] = (np.arange(16).reshape(4, 4) % 2) * 900  # block 0 difficult (uint32 wrap quirk)
    cases.append(("block0_difficult", b0.astype(np.uint16), True))
    flat = np.full((32, 32), 1000, dtype=np.uint16)
    cases.append(("flat", flat, True))
    cases.append(("zeros", np.zeros((64, 64), dtype=np.uint16), True))
    d64 = np.zeros((16, 16), dtype=np.uint16)
    d64.flat[::2] = 64  # alternating +64 / -64: the |d|>64 vs -64<d<65 mismatch
    cases.append(("delta_pm64", d64, True))
    cases.append(("synth_ct_64", synth.synth_ct(5, 64), True))
    cases.append(("synth_ct_128", synth.synth_ct(6, 128), True))
    for name, im, in_domain in cases:
        small[f"{name}__in"] = im
        small[f"{name}__out"] = np.frombuffer(ref_encode(im), dtype=np.uint8)
        if in_domain:
            assert (ref_decode(small[f"{name}__out"].tobytes(), im.shape) == im).all(), name
        small[f"{name}__domain"] = np.array(in_domain)
    # all 8 flag variants on two images
    for name in ("walk_32x64", "synth_ct_64"):
        im = small[f"{name}__in"]
        for fr in (0, 1):
            for sg in (0, 1):
                for df in (0, 1):
                    out = ref_encode(im, fractal=bool(fr), seg=bool(sg), deflate=bool(df))
                    small[f"{name}__flags{fr}{sg}{df}"] = np.frombuffer(out, dtype=np.uint8)
                    assert (ref_decode(out, im.shape) == im).all()
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **small)

    # 3. full-size slices: digests only
    ct = {}
    for seed in (0, 1, 2, 3, 4095):
        im = synth.synth_ct(seed)
        out = ref_encode(im)
        assert (ref_decode(out, im.shape) == im).all()
        ct[str(seed)] = {"sha256": hashlib.sha256(out).hexdigest(), "len": len(out),
                         "input_sha256": hashlib.sha256(im.tobytes()).hexdigest()}
        print("synth_ct", seed, len(out), flush=True)
    digests["synth_ct_512"] = ct
    im = synth.synth_radiograph(0, 2048)
    out = ref_encode(im)
    assert (ref_decode(out, im.shape) == im).all()
    digests["synth_radiograph_2048_seed0"] = {"sha256": hashlib.sha256(out).hexdigest(), "len": len(out),
                                              "input_sha256": hashlib.sha256(im.tobytes()).hexdigest()}
    print("radiograph", len(out), flush=True)
    import zlib
    digests["zlib_runtime_version"] = zlib.ZLIB_RUNTIME_VERSION
    digests["numpy_version"] = np.__version__
    json.dump(digests, open(os.path.join(HERE, "golden_digests.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
