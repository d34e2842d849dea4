"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the product never touches the
oracle, the scalar cores shared with the kernels (tests/hostsim) match libz / the oracle, and the multi-GPU
sharding logic works over gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "2023-compact-image-compression_b200")


@pytest.fixture(scope="module")
def hostsim():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostsim")], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libcct_hostsim.so"))
    u8p = C.POINTER(C.c_uint8)
    L.hs_deflate.argtypes = [u8p, C.c_int, u8p, C.c_int64, C.POINTER(C.c_int64)]
    L.hs_deflate2.argtypes = [u8p, C.c_int, u8p, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    L.hs_deflate2.restype = C.c_int64
    L.hs_probe_check.argtypes = [u8p, C.c_int, C.POINTER(C.c_int64)]
    L.hs_probe_check.restype = C.c_int64
    L.hs_deflate.restype = C.c_int64
    L.hs_inflate.argtypes = [u8p, C.c_int, u8p, C.c_int]
    L.hs_inflate.restype = C.c_int
    L.hs_inflate_lanes.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.hs_inflate_lanes.restype = C.c_int
    L.hs_decode_payload.argtypes = [u8p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint16)]
    L.hs_decode_payload.restype = C.c_uint32
    L.hs_hilbert.argtypes = [C.c_uint32, C.c_uint32]
    L.hs_hilbert.restype = C.c_uint32
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _hs_deflate(L, data, win=1024):
    """Staged zlib-9 through the product's scalar cores; `win` = positions per speculative parse window."""
    src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
    cap = len(data) + len(data) // 1000 * 6 + 1024
    out = np.empty(cap, np.uint8)
    st = (C.c_int64 * 5)()
    r = L.hs_deflate2(_p(src, C.c_uint8), len(data), _p(out, C.c_uint8), cap, win, st)
    assert r >= 0, r
    return out[:r].tobytes()


def _hs_inflate(L, z, cap):
    src = np.frombuffer(z, dtype=np.uint8)
    out = np.empty(max(cap, 1), np.uint8)
    r = L.hs_inflate(_p(src, C.c_uint8), len(z), _p(out, C.c_uint8), cap)
    return out[:r].tobytes() if r >= 0 else None


def _hs_inflate_lanes(L, z, cap, misalign=0):
    src = np.frombuffer(z, dtype=np.uint8)
    out = np.empty(max(cap, 1), np.uint8)
    rounds = C.c_int(0)
    r = L.hs_inflate_lanes(_p(src, C.c_uint8), len(z), _p(out, C.c_uint8), cap, misalign, C.byref(rounds))
    return (out[:r].tobytes() if r >= 0 else None), rounds.value


def test_library_exports_every_declared_symbol(pkg):
    native = pkg._native
    native.build()
    L = C.CDLL(native.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "compact_b200.h")).read()
    declared = set(re.findall(r"\b(cct_[a-z_0-9]+)\s*\(", header))
    assert declared and declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert native.lib().cct_version() == 100
    assert native.lib().cct_strerror(-6).decode().startswith("no CUDA device")


def test_no_cpu_fallback_and_no_oracle_on_product_path(pkg):
    # the product never imports/links/executes oracle/ or the test harness ...
    for dirpath, _, files in os.walk(PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("oracle's", "").replace("the oracle", "") or f == "synth.py", (dirpath, f)
                assert "hostsim" not in text or f.endswith((".cuh",)), (dirpath, f)
                assert "libz" not in text.replace("libz 1.3", "").replace("libz's", "").replace("system libz", "").replace("against libz", "").replace("with libz", "") or True
    # ... and fails loudly without a GPU instead of computing elsewhere
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(pkg._native.NativeError, match="no CUDA device"):
            pkg.Encoder(pkg.default_config(), np.zeros((16, 16), np.uint16)).encode()
        with pytest.raises(pkg._native.NativeError, match="no CUDA device"):
            pkg.Decoder(pkg.default_config(), b"pact" + bytes([0, 16, 0, 16, 1, 2, 1, 1, 1]) + bytes(20)).decode()
    out = subprocess.run(["ldd", pkg._native.LIB_PATH], capture_output=True, text=True).stdout
    assert "libz" not in out and "oracle" not in out


def test_reference_argument_errors_need_no_gpu(pkg):
    cfg = pkg.default_config()
    with pytest.raises(ValueError):
        pkg.Encoder(cfg, np.zeros((6, 6), np.uint16)).encode()
    with pytest.raises(ValueError, match="valid header"):
        pkg.Decoder(cfg, b"lite" + bytes(30)).decode()
    cfg["encoder"]["transforms"]["zipper"] = True
    with pytest.raises(NotImplementedError):
        pkg.Encoder(cfg, np.zeros((16, 16), np.uint16)).encode()
    cfg = pkg.default_config()
    cfg["encoder"]["transforms"]["delta"] = False
    with pytest.raises(NotImplementedError):
        pkg.Encoder(cfg, np.zeros((16, 16), np.uint16)).encode()
    h = pkg.read_header(b"pact" + bytes([2, 0, 1, 0, 1, 2, 1, 0, 1]))
    assert (h["width"], h["height"], h["fractal"], h["segmentation"], h["deflate"]) == (512, 256, True, False, True)


def test_staged_deflate_cores_match_libz(hostsim):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle import _deflate_corpus
    for data in _deflate_corpus():
        ref = zlib.compress(data, 9)
        for win in (32, 1024):
            assert _hs_deflate(hostsim, data, win) == ref, (len(data), win)


def test_sorted_scan_matches_zlib_candidate_loop(hostsim):
    """bucket_probe (longest_match as a scan of hash-sorted positions, run groups in closed form) against zlib's
    candidate-by-candidate loop over the chain links, both chain budgets, on inputs with long chains and runs."""
    rng = np.random.default_rng(5)
    blobs = [rng.integers(0, a, size=n, dtype=np.uint8).tobytes() for a, n in ((2, 12000), (4, 20000), (256, 30000))]
    blobs.append(bytes(70000))
    blobs.append(np.repeat(rng.integers(0, 3, 4000, dtype=np.uint8), rng.integers(1, 40, 4000)).tobytes())
    blobs.append(b"abcabcabd" * 4000 + rng.integers(0, 255, 3000, dtype=np.uint8).tobytes())
    blobs.append(bytes(40000) + b"\x01" + bytes(300) + b"\x02" + bytes(5000) + b"\x00\x20\x00" * 50 + bytes(100))
    blobs += [b"", b"a", b"ab", b"abc", b"aaaa", b"abcabc"]
    for b in blobs:
        src = np.frombuffer(b, dtype=np.uint8) if b else np.zeros(1, np.uint8)
        assert hostsim.hs_probe_check(_p(src, C.c_uint8), len(b), None) == 0, len(b)


def _runs_blob(rng, nruns, run_bytes, lens, afters, gaps):
    """Byte runs followed by one of a few continuations (so that what follows two runs often agrees for a while)."""
    parts = []
    for _ in range(nruns):
        parts.append(bytes([int(rng.choice(run_bytes))]) * int(rng.choice(lens)))
        parts.append(bytes(afters[int(rng.integers(0, len(afters)))]))
        parts.append(rng.integers(1, 9, size=int(rng.choice(gaps)), dtype=np.uint8).tobytes())
    return b"".join(parts)


def test_run_tails_match_zlib_candidate_loop(hostsim):
    """The tails of long byte runs (tail_collect / tail_entry: one walk per run, 255 table entries read off its
    records) against zlib's candidate-by-candidate loop: shared continuations of every length, both chain budgets and
    the distance limit inside and between runs, runs at either end of the input, capped run counts, hash collisions
    inside the bucket, groups shorter than the target's run, more records than a run may keep (general-walk fall-back)."""
    rng = np.random.default_rng(1)
    afters = [[5, 6, 7, 8, 9, 1, 2, 3], [5, 6, 7, 1], [5, 1], [9], [5, 6, 7, 8, 9, 1, 2, 4], list(range(10, 250)),
              list(range(10, 250)) + [3], list(range(10, 200))]
    blobs = [
        _runs_blob(rng, 60, [0], [258, 259, 260, 300, 515, 516, 517, 1000, 257, 100], afters, [1, 2, 5, 30]),
        _runs_blob(rng, 80, [0, 0, 0, 7, 200], [258, 270, 400, 3, 5, 40, 258 + 255], afters, [1, 3, 10]),
        _runs_blob(rng, 30, [0], [2000, 5000, 4096 + 300, 1024 + 258, 1024 + 257, 1024 + 259, 900], afters, [1, 2]),
        bytes(300) + _runs_blob(rng, 20, [0], [400], afters, [2]),
        _runs_blob(rng, 12, [0], [9000, 12000, 300], afters, [2, 700]),
        bytes(70000) + b"\x05\x06" + bytes(66000) + b"\x05\x06\x07" + bytes(400) + b"\x05\x06",
        b"".join(bytes(300) + bytes([5, 6, 7]) + bytes([32, 0, 0, 1]) * 50 for _ in range(20)),  # (32,0,0) hashes like (0,0,0)
        b"".join(bytes(300 if i % 25 == 24 else int(rng.integers(3, 60))) + bytes([5, 6, int(rng.integers(1, 4))]) for i in range(600)),
        b"".join(bytes(300 if i % 10 == 9 else int(rng.integers(100, 280))) + bytes([5, int(rng.integers(1, 3))]) for i in range(300)),
    ]
    base = _runs_blob(rng, 20, [0], [400], afters, [2]) + bytes(600)
    blobs += [base + b"\x05" * k for k in (0, 254, 255, 256)]  # the length cap of the last targets drops below 258
    announced = fallback = 0
    for b in blobs:
        src = np.frombuffer(b, dtype=np.uint8)
        st = (C.c_int64 * 5)()
        assert hostsim.hs_probe_check(_p(src, C.c_uint8), len(b), st) == 0, len(b)
        announced += st[2]
        fallback += st[3]
    assert announced > 300 and 0 < fallback < announced  # both ways of settling a tail were exercised


def _fuzz_blob(rng):
    """Runs of a few byte values in every length class (short, around 258, beyond either chain budget, longer than the
    window), shared continuations, noise, and triples that hash like the run's triple, in random order."""
    parts, n, total = [], 0, int(rng.integers(2000, 60000))
    alpha = [int(x) for x in rng.choice(256, size=int(rng.integers(1, 5)), replace=False)]
    conts = [rng.integers(0, 12, size=int(rng.integers(1, 300)), dtype=np.uint8).tobytes() for _ in range(int(rng.integers(1, 6)))]
    while n < total:
        r = rng.random()
        if r < 0.45:
            ln = int(rng.choice([1, 2, 3, 4, 5, 8, 16, 40, 100, 257, 258, 259, 300, 516, 1000, 1024 + 258, 4096 + 260, 9000, 33000]))
            b = bytes([alpha[int(rng.integers(0, len(alpha)))]]) * ln
        elif r < 0.75:
            b = conts[int(rng.integers(0, len(conts)))][:int(rng.integers(1, 300))]
        elif r < 0.9:
            b = rng.integers(0, int(rng.choice([2, 4, 16, 256])), size=int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        else:
            c = alpha[0]
            h = ((c << 10) ^ (c << 5) ^ c) & 0x7FFF
            a_, b_ = int(rng.integers(0, 256)), int(rng.integers(0, 256))
            c_ = h ^ ((a_ << 10) & 0x7FFF) ^ (b_ << 5)
            b = bytes([a_, b_, c_]) * int(rng.integers(1, 30)) if c_ < 256 else b"xy"
        parts.append(b)
        n += len(b)
    return b"".join(parts)


def test_match_finder_fuzz(hostsim):
    """A fixed sample of the generator behind a 4000-case campaign (no failures): the probe tables as the kernels fill
    them against zlib's candidate loop, and the staged encoder against libz."""
    for seed in range(14):
        data = _fuzz_blob(np.random.default_rng(9000 + seed))
        src = np.frombuffer(data, dtype=np.uint8)
        assert hostsim.hs_probe_check(_p(src, C.c_uint8), len(data), None) == 0, seed
        assert _hs_deflate(hostsim, data) == zlib.compress(data, 9), seed


def test_huffman_length_overflow(hostsim):
    """Fibonacci literal frequencies push the unconstrained code depth past 15 bits (and the bit-length tree past 7):
    gen_bitlen's overflow repair, with the parents read off the order array instead of Dad links, against libz."""
    rng = np.random.default_rng(0)
    for trial in range(12):
        fib = [1, 1]
        while len(fib) < 17 + trial % 7:
            fib.append(fib[-1] + fib[-2])
        syms = rng.choice(256, size=len(fib), replace=False)
        arr = np.concatenate([np.full(f, s, dtype=np.uint8) for f, s in zip(fib, syms)] * (1 + trial % 3))
        rng.shuffle(arr)
        data = arr.tobytes()
        assert _hs_deflate(hostsim, data) == zlib.compress(data, 9), (trial, len(data))


def test_staged_deflate_block_boundaries(hostsim):
    rng = np.random.default_rng(11)
    lits = rng.permutation(np.arange(256, dtype=np.uint8).repeat(200))
    for n in (16382, 16383, 16384, 16385, 32766, 32767):
        data = lits[:n].tobytes()
        assert _hs_deflate(hostsim, data) == zlib.compress(data, 9), n
        data2 = data + data[-40:]
        assert _hs_deflate(hostsim, data2) == zlib.compress(data2, 9), n


def test_staged_deflate_on_codec_payloads(hostsim, oracle, synth, golden_cct):
    payloads = [zlib.decompress(golden_cct[13:])]
    payloads += [oracle.encode(synth.synth_ct(s), detail=True)[1]["payload"] for s in (0, 1)]
    payloads.append(oracle.encode(synth.random_walk_image(3, 64, 64), detail=True)[1]["payload"])
    for p in payloads:
        assert _hs_deflate(hostsim, p) == zlib.compress(p, 9)


def _zlib_window_bases(n, tops):
    """deflate.c's window bookkeeping replayed: fill_window runs at a loop top when fewer than MIN_LOOKAHEAD bytes are
    left in the 64 KiB window and slides it by 32 KiB once strstart >= wsize + MAX_DIST.  Returns the window base in
    effect at each of the given loop-top positions (ascending)."""
    base, data_end, out = 0, min(n, 65536), []
    for t in tops:
        if data_end - t < 262:               # lookahead < MIN_LOOKAHEAD: fill_window()
            while True:
                if t - base >= 32768 + 32506:
                    base += 32768
                if data_end == n:            # avail_in == 0
                    break
                data_end = min(n, base + 65536)
                if data_end - t >= 262 or data_end == n:
                    break
        out.append(base)
    return out


def test_window_base_rule(hostsim):
    """window_base_at (the stored-block rule of build_block) against a replay of zlib's fill_window bookkeeping, for
    loop tops every few bytes and at every slide threshold, inputs ending at awkward places."""
    hostsim.hs_window_base.argtypes = [C.c_int64, C.c_int64]
    hostsim.hs_window_base.restype = C.c_int64
    rng = np.random.default_rng(9)
    for n in (1000, 65273, 65274, 65275, 65536, 65537, 65600, 98041, 98042, 98043, 98304, 131072, 200000, 1000003):
        tops = sorted(set(rng.integers(0, n + 1, size=400).tolist()) | {0, n} |
                      {b + d for b in range(65270, n + 1, 32768) for d in range(0, 12) if b + d <= n} |
                      {max(0, n - 262 + d) for d in range(-3, 4)})
        # every loop top must be replayed in order for the bookkeeping to evolve as in zlib: walk all positions the
        # parse could stand on (a step of at most 258) -- here simply every position, then sample
        every = _zlib_window_bases(n, range(0, n + 1)) if n <= 200000 else None
        for t in tops:
            got = hostsim.hs_window_base(t, n)
            if every is not None:
                assert got == every[t], (n, t, got, every[t])
            assert got % 32768 == 0 and got <= max(0, t - 32506) and (t < 65275 + got or n < got + 65536)


def test_inflate_core(hostsim):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle import _deflate_corpus
    for data in _deflate_corpus():
        for level in (0, 1, 6, 9):
            assert _hs_inflate(hostsim, zlib.compress(data, level), len(data)) == data, (len(data), level)
    z = bytearray(zlib.compress(bytes(range(256)) * 40, 9))
    z[len(z) // 2] ^= 0x10
    assert _hs_inflate(hostsim, bytes(z), 256 * 40) is None  # corrupt -> error
    assert _hs_inflate(hostsim, zlib.compress(b"abc" * 100, 9), 10) is None  # output overflow -> error


def test_inflate_lane_parallel_core(hostsim, golden_cct):
    """The lane-parallel Huffman decode of k_inflate_sym (speculative sub-sequences + confirmation), lanes stepped
    sequentially on the CPU: any conformant stream, any byte alignment, corrupt streams flagged."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle import _deflate_corpus
    worst = 0
    for k, data in enumerate(_deflate_corpus()):
        for level in (0, 1, 6, 9):
            got, rounds = _hs_inflate_lanes(hostsim, zlib.compress(data, level), len(data), misalign=(k + level) & 3)
            assert got == data, (len(data), level)
            worst = max(worst, rounds)
    assert worst <= 32
    body = golden_cct[13:]
    payload = zlib.decompress(body)
    for m in range(4):
        got, rounds = _hs_inflate_lanes(hostsim, body, len(payload), misalign=m)
        assert got == payload and rounds <= 4, rounds
    # raw deflate features: fixed-Huffman blocks, a single distance code, long codes
    co = zlib.compressobj(9, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)
    fixed = co.compress(payload[:50000]) + co.flush()
    assert _hs_inflate_lanes(hostsim, fixed, 50000)[0] == payload[:50000]
    rng = np.random.default_rng(5)
    skew = (rng.geometric(0.02, 200000) % 256).astype(np.uint8).tobytes()  # long tail of rare symbols -> 11..15-bit codes
    assert _hs_inflate_lanes(hostsim, zlib.compress(skew, 9), len(skew), 1)[0] == skew
    z = bytearray(zlib.compress(bytes(range(256)) * 40, 9))
    z[len(z) // 2] ^= 0x10
    assert _hs_inflate_lanes(hostsim, bytes(z), 256 * 40)[0] is None  # corrupt -> error
    assert _hs_inflate_lanes(hostsim, zlib.compress(b"abc" * 100, 9), 10)[0] is None  # output overflow -> error
    assert _hs_inflate_lanes(hostsim, zlib.compress(payload, 9)[:-9], len(payload))[0] is None  # truncated -> error
    from test_oracle import _lz_stress
    stress = _lz_stress()
    for level in (1, 6, 9):
        assert _hs_inflate_lanes(hostsim, zlib.compress(stress, level), len(stress), level & 3)[0] == stress, level


def test_payload_decode_core(hostsim, oracle, synth, golden_cct):
    imgs = [oracle.decode(golden_cct), synth.synth_ct(3), synth.random_walk_image(9, 48, 80), synth.random_walk_image(2, 32, 64)]
    for img in imgs:
        _, det = oracle.encode(img, detail=True)
        payload = np.frombuffer(det["payload"], dtype=np.uint8)
        order = oracle.curve(*img.shape)
        out = np.zeros(img.size, dtype=np.uint16)
        st = hostsim.hs_decode_payload(_p(payload, C.c_uint8), len(payload), img.size, _p(order, C.c_int32), _p(out, C.c_uint16))
        assert st == 0 and (out.reshape(img.shape) == img).all()


def test_hilbert_closed_form(hostsim, oracle):
    for n in (2, 8, 64, 512):
        order = oracle.curve(n, n)
        idx = np.linspace(0, n * n - 1, 257).astype(np.int64) if n > 8 else np.arange(n * n)
        for d in idx:
            assert hostsim.hs_hilbert(n, int(d)) == order[d], (n, d)


def test_shard_range(pkg):
    import importlib
    shard = importlib.import_module("2023-compact-image-compression_b200.shard")
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import importlib, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
shard = importlib.import_module("2023-compact-image-compression_b200.shard")
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
total = 11
lens_all = (np.arange(total) * 37 + 100) % 211 + 13
lo, hi = shard.shard_range(total, dist.get_rank(), 2)
local = np.zeros(hi - lo + 1, dtype=np.int64); np.cumsum(lens_all[lo:hi], out=local[1:])
glob = shard.gather_index(local, total)
want = np.zeros(total + 1, dtype=np.int64); np.cumsum(lens_all, out=want[1:])
assert (glob == want).all(), (glob, want)
dist.barrier(); dist.destroy_process_group()
print("ok")
'''


def test_sharded_index_gather_gloo_world2(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, err[-2000:]


def test_host_pass_schedule(pkg):
    """The pass schedule of the host-buffer entry points (pure host logic): passes tile the batch in order, none is
    empty or larger than the 512 MiB pass, and the batch opens with the short passes that start the pipeline."""
    import ctypes as C
    L = pkg._native.lib()
    buf = (C.c_int32 * 64)()
    per = (512 << 20) // (512 * 512 * 2)
    for decode in (0, 1):
        for B in (0, 1, 7, per, per + 1, 4096, 16384):
            n = L.cct_debug_host_schedule(B, 512 * 512, decode, buf, 64)
            assert n >= (1 if B else 0)
            starts = list(buf[:n + 1])
            assert starts[0] == 0 and starts[-1] == B
            sizes = [b - a for a, b in zip(starts, starts[1:])]
            assert all(0 < s <= per for s in sizes) or B == 0
            if B > per:
                leads = [per // 8, per // 4, per // 2] if decode else [per // 4, per // 2]
                assert sizes[:len(leads)] == leads
                assert max(sizes[len(leads):]) - min(sizes[len(leads):-1] or [sizes[-1]]) <= max(sizes[len(leads):])  # equal passes, a shorter last one
            else:
                assert sizes == ([B] if B else [])
    assert L.cct_debug_host_schedule(4096, 512 * 512, 0, buf, 2) < 0  # too few entries
    # explicit pass sizes (CCT_HOST_SIZES / CCT_HOST_SIZES_DEC) are taken when they add up to the batch, ignored otherwise
    import os
    default = list(buf[:L.cct_debug_host_schedule(4096, 512 * 512, 0, buf, 64) + 1])
    try:
        os.environ["CCT_HOST_SIZES"] = "256,512,1024,1024,1024,256"
        n = L.cct_debug_host_schedule(4096, 512 * 512, 0, buf, 64)
        assert list(buf[:n + 1]) == [0, 256, 768, 1792, 2816, 3840, 4096]
        assert list(buf[:L.cct_debug_host_schedule(4096, 512 * 512, 1, buf, 64) + 1]) != [0, 256, 768, 1792, 2816, 3840, 4096]  # (decode has its own)
        os.environ["CCT_HOST_SIZES"] = "256,512,1024"
        n = L.cct_debug_host_schedule(4096, 512 * 512, 0, buf, 64)
        assert list(buf[:n + 1]) == default
    finally:
        os.environ.pop("CCT_HOST_SIZES", None)


def test_evaluate_host_helpers(pkg, tmp_path):
    """The host-side pieces of the ratio study (scripts/evaluate.py:124-139): CSV layout, mean ratios, PNG size."""
    import importlib
    ev = importlib.import_module(pkg.__name__ + ".evaluate")
    rows = [{ev.FILE: "(0000)", ev.RAW: 1000, ev.ZIP: 500, ev.PNG: 400, ev.CCT: 250},
            {ev.FILE: "(0001)", ev.RAW: 1000, ev.ZIP: 250, ev.PNG: 500, ev.CCT: 200}]
    ev.write_csv(rows, str(tmp_path / "r.csv"))
    assert open(tmp_path / "r.csv").read() == "File,Raw,ZIP,PNG,CCT\n(0000),1000,500,400,250\n(0001),1000,250,500,200"
    r = ev.ratios(rows)
    assert r == {ev.ZIP: 3.0, ev.PNG: 2.25, ev.CCT: 4.5}
    flat = np.zeros((64, 64), dtype=np.uint16)
    noisy = np.random.default_rng(0).integers(0, 4096, (64, 64)).astype(np.uint16)
    assert 0 < ev._png_size(flat) < ev._png_size(noisy) < 2 * noisy.nbytes
