"""Host-buffer path: encode / decode time of the 4096-slice batch for explicit pass sizes (CCT_HOST_SIZES[_DEC]).
Run on the GPU box from the repository root: python profiles/src/e2e_sweep.py"""
import importlib, os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("2023-compact-image-compression_b200")
synth = importlib.import_module("2023-compact-image-compression_b200.synth")
native = pkg._native; L = native.lib()
B, n = 4096, 512
base = synth.synth_ct_batch(range(512), workers=16)
imgs = np.concatenate([base] * (B // 512))
h_px = torch.from_numpy(imgs.view(np.int16)).pin_memory()
cap = B * int(L.cct_file_bound(n, n))
h_arena = torch.empty(cap // 3, dtype=torch.uint8).pin_memory()
h_off = torch.zeros(B + 1, dtype=torch.int64).pin_memory()
h_st = torch.zeros(B, dtype=torch.int32).pin_memory()
h_out = torch.empty((B, n, n), dtype=torch.int16).pin_memory()
def enc(): native.check(L.cct_encode_host(h_px.data_ptr(), B, n, n, 7, h_arena.data_ptr(), h_arena.numel(), h_off.data_ptr(), h_st.data_ptr(), None, 0))
def dec(): native.check(L.cct_decode_host(h_arena.data_ptr(), h_off.data_ptr(), B, n, n, h_out.data_ptr(), h_st.data_ptr(), 0))
def t(fn, reps=5):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
enc(); dec(); assert torch.equal(h_out, h_px)
enc_cfgs = ["", "256,512,768,768,768,512,256,256", "256,512,896,896,896,384,256", "256,512,1024,1024,768,512",
            "128,256,512,896,896,768,384,256", "256,512,704,704,704,704,512", "256,384,640,768,768,640,384,256",
            "192,320,512,768,768,640,512,384", "256,512,832,832,832,416,416", "256,512,1024,1024,640,384,256",
            "256,512,960,960,704,448,256", "128,384,768,960,960,512,256,128", ""]
for c in enc_cfgs:
    os.environ.pop("CCT_HOST_SIZES", None)
    if c: os.environ["CCT_HOST_SIZES"] = c
    assert sum(int(x) for x in c.split(",")) == B if c else True
    print("encode %-40s %.2f ms" % (c or "default", t(enc)), flush=True)
os.environ.pop("CCT_HOST_SIZES", None)
dec_cfgs = ["", "64,128,256,512,1024,1024,1088", "128,256,512,1024,1024,1152", "32,96,256,512,1024,1024,1152",
            "128,128,256,512,768,768,768,768", "64,64,128,256,512,1024,1024,1024", ""]
for c in dec_cfgs:
    os.environ.pop("CCT_HOST_SIZES_DEC", None)
    if c: os.environ["CCT_HOST_SIZES_DEC"] = c
    assert sum(int(x) for x in c.split(",")) == B if c else True
    print("decode %-40s %.2f ms" % (c or "default", t(dec)), flush=True)
