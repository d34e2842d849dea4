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
def run(kind, fn, cfgs):
    keys = ("CCT_HOST_SIZES", "CCT_HOST_SIZES_DEC", "CCT_HOST_LANES", "CCT_HOST_PASS", "CCT_HOST_LEADS", "CCT_HOST_TAILS")
    for env in cfgs:
        for k in keys: os.environ.pop(k, None)
        os.environ.update(env)
        print("%s %-60s %.2f ms" % (kind, " ".join(f"{k[9:]}={v}" for k, v in env.items()) or "default", t(fn)), flush=True)
    for k in keys: os.environ.pop(k, None)
L3 = {"CCT_HOST_LANES": "3"}
run("encode", enc, [{}, L3, {**L3, "CCT_HOST_PASS": "768"}, {**L3, "CCT_HOST_PASS": "640"}, {**L3, "CCT_HOST_PASS": "512"},
                    {**L3, "CCT_HOST_PASS": "768", "CCT_HOST_LEADS": "4,2,2", "CCT_HOST_TAILS": "2,2,4"},
                    {**L3, "CCT_HOST_SIZES": "256,384,512,640,640,640,512,320,192"},
                    {**L3, "CCT_HOST_SIZES": "192,320,448,704,704,704,512,320,192"},
                    {"CCT_HOST_PASS": "768"}, {}])
run("decode", dec, [{}, L3, {**L3, "CCT_HOST_PASS": "768"}, {**L3, "CCT_HOST_PASS": "640"}, {"CCT_HOST_LANES": "1"}, {}])
