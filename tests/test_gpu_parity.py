"""GPU parity tests: the CUDA path (through the C ABI / its Python mirror) against the oracle and the
golden fixtures.  Bit-exact everywhere (integer/byte work).  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import hashlib
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def native(pkg):
    L = pkg._native.lib()
    assert L.cct_device_count() >= 1
    return pkg._native


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _stages(torch, native, img, flags=7):
    """cct_debug_encode_stages for one image -> (order, jump, payload bytes)."""
    L = native.lib()
    rows, cols = img.shape
    n = rows * cols
    d_px = _dev(torch, img.view(np.int16))
    ws = torch.empty(int(L.cct_encode_workspace_bytes(1, rows, cols, flags)), dtype=torch.uint8, device="cuda")
    d_order = torch.empty(n, dtype=torch.int32, device="cuda")
    d_jump = torch.empty(n // 16, dtype=torch.uint8, device="cuda")
    d_payload = torch.empty(int(L.cct_payload_bound(rows, cols)), dtype=torch.uint8, device="cuda")
    d_plen = torch.zeros(1, dtype=torch.int32, device="cuda")
    native.check(L.cct_debug_encode_stages(d_px.data_ptr(), rows, cols, flags, d_order.data_ptr(), d_jump.data_ptr(),
                                           d_payload.data_ptr(), d_plen.data_ptr(), ws.data_ptr(), ws.numel(),
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    plen = int(d_plen.item())
    return d_order.cpu().numpy(), d_jump.cpu().numpy(), d_payload[:plen].cpu().numpy().tobytes()


def _gpu_deflate(torch, native, blobs):
    L = native.lib()
    B = len(blobs)
    max_len = max(max(len(b) for b in blobs), 1)
    in_stride = (max_len + 15) // 16 * 16
    out_stride = int(L.cct_deflate_bound(max_len))
    h_in = np.zeros((B, in_stride), dtype=np.uint8)
    for i, b in enumerate(blobs):
        h_in[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    d_in = _dev(torch, h_in)
    d_len = _dev(torch, np.array([len(b) for b in blobs], dtype=np.int32))
    d_out = torch.empty((B, out_stride), dtype=torch.uint8, device="cuda")
    d_olen = torch.zeros(B, dtype=torch.int32, device="cuda")
    ws = torch.empty(int(L.cct_deflate_workspace_bytes(B, max_len)), dtype=torch.uint8, device="cuda")
    native.check(L.cct_deflate_batch(d_in.data_ptr(), in_stride, d_len.data_ptr(), B, max_len, d_out.data_ptr(),
                                     out_stride, d_olen.data_ptr(), ws.data_ptr(), ws.numel(),
                                     torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out, olen = d_out.cpu().numpy(), d_olen.cpu().numpy()
    return [out[i, :olen[i]].tobytes() for i in range(B)]


def _gpu_inflate(torch, native, streams, cap):
    L = native.lib()
    B = len(streams)
    in_stride = max(len(s) for s in streams) + 16
    h_in = np.zeros((B, in_stride), dtype=np.uint8)
    for i, s in enumerate(streams):
        h_in[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    d_in = _dev(torch, h_in)
    d_len = _dev(torch, np.array([len(s) for s in streams], dtype=np.int32))
    d_out = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
    d_olen = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_st = torch.zeros(B, dtype=torch.int32, device="cuda")
    ws = torch.empty(int(L.cct_inflate_workspace_bytes(B, cap)), dtype=torch.uint8, device="cuda")
    native.check(L.cct_inflate_batch(d_in.data_ptr(), in_stride, d_len.data_ptr(), B, d_out.data_ptr(), cap,
                                     d_olen.data_ptr(), d_st.data_ptr(), ws.data_ptr(), ws.numel(),
                                     torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out, olen, st = d_out.cpu().numpy(), d_olen.cpu().numpy(), d_st.cpu().numpy()
    return [out[i, :olen[i]].tobytes() for i in range(B)], st


# ------------------------------------------------------------------------------------------ stage parity

@pytest.mark.parametrize("shape", [(512, 512), (64, 64), (32, 64), (64, 32), (48, 80), (20, 20), (12, 12), (4, 4), (1, 16), (36, 44)])
def test_traversal_table(torch_cuda, native, oracle, synth, shape):
    img = synth.random_walk_image(1, *shape)
    order, _, _ = _stages(torch_cuda, native, img)
    assert (order == oracle.curve(*shape)).all()


def test_stage_payload_and_jumps(torch_cuda, native, oracle, synth, golden_cct):
    imgs = [synth.synth_ct(s) for s in (0, 1, 2)] + [oracle.decode(golden_cct)]
    imgs += [synth.random_walk_image(50 + k, h, w) for k, (h, w) in enumerate([(64, 64), (32, 64), (48, 80), (128, 128), (16, 16)])]
    b0 = synth.random_walk_image(7, 32, 32)
    b0[:4, :4] = (np.arange(16).reshape(4, 4) % 2) * 900  # difficult block 0: uint32-wrap quirk
    imgs.append(b0)
    for img in imgs:
        for flags in (7, 5, 6, 4):  # with / without fractal and segmentation
            _, det = oracle.encode(img, fractal=bool(flags & 1), seg=bool(flags & 2), detail=True)
            _, jump, payload = _stages(torch_cuda, native, img, flags)
            want_jump = np.where(det["jumps"] >= 0, det["jumps"] - np.arange(len(det["jumps"])), 0).astype(np.uint8)
            assert (jump == want_jump).all(), (img.shape, flags)
            assert payload == det["payload"], (img.shape, flags)


def _corpus():
    from test_oracle import _deflate_corpus
    return _deflate_corpus()


def test_deflate_matches_libz(torch_cuda, native, oracle, synth):
    blobs = [b for b in _corpus() if len(b) > 0]
    for s in (0, 1):
        blobs.append(oracle.encode(synth.synth_ct(s), detail=True)[1]["payload"])
    from test_oracle import _lz_stress
    blobs += [_lz_stress(3, 300000), _lz_stress(4, 70001)]
    from test_host_logic import _runs_blob
    rng_r = np.random.default_rng(12)
    afters = [[5, 6, 7, 8, 9, 1, 2, 3], [5, 6, 7, 1], [5, 1], [9], list(range(10, 250))]
    blobs += [_runs_blob(rng_r, 200, [0, 0, 7], [258, 300, 517, 1000, 40, 5, 3, 2000, 5000], afters, [1, 2, 5, 30, 300]),
              b"".join(bytes(300 if i % 25 == 24 else int(rng_r.integers(3, 60))) + bytes([5, 6, int(rng_r.integers(1, 4))]) for i in range(600))]
    for trial in range(6):  # Fibonacci literal frequencies: code depths beyond 15 bits, gen_bitlen's overflow repair
        fib = [1, 1]
        while len(fib) < 17 + trial:
            fib.append(fib[-1] + fib[-2])
        arr = np.concatenate([np.full(f, s, dtype=np.uint8) for f, s in zip(fib, rng_r.choice(256, size=len(fib), replace=False))] * (1 + trial % 3))
        rng_r.shuffle(arr)
        blobs.append(arr.tobytes())
    got = _gpu_deflate(torch_cuda, native, blobs)
    for b, g in zip(blobs, got):
        assert g == zlib.compress(b, 9), len(b)


def test_deflate_block_boundaries(torch_cuda, native):
    rng = np.random.default_rng(11)
    lits = rng.permutation(np.arange(256, dtype=np.uint8).repeat(200))
    blobs = []
    for n in (16382, 16383, 16384, 16385, 32765, 32766, 32767, 32768):
        data = lits[:n].tobytes()
        blobs += [data, data + data[-40:]]
    got = _gpu_deflate(torch_cuda, native, blobs)
    for b, g in zip(blobs, got):
        assert g == zlib.compress(b, 9), len(b)


def test_deflate_empty_and_tiny(torch_cuda, native):
    blobs = [b"a", b"ab", b"abc", b"aaaa", bytes(5), bytes(300)]
    got = _gpu_deflate(torch_cuda, native, blobs)
    for b, g in zip(blobs, got):
        assert g == zlib.compress(b, 9), b[:8]


def test_inflate_any_conformant_stream(torch_cuda, native):
    blobs = [b for b in _corpus() if len(b) > 0]
    cap = max(len(b) for b in blobs)
    for level in (0, 1, 6, 9):
        streams = [zlib.compress(b, level) for b in blobs]
        got, st = _gpu_inflate(torch_cuda, native, streams, cap)
        assert not st.any()
        for b, g in zip(blobs, got):
            assert g == b, (level, len(b))
    # copies from anywhere in the window, byte runs and literal bursts mixed; rows that do not start on 16 bytes
    from test_oracle import _lz_stress
    stress = [_lz_stress(3, 300000), _lz_stress(4, 70001), _lz_stress(5, 123457)]
    for level in (1, 6, 9):
        got, st = _gpu_inflate(torch_cuda, native, [zlib.compress(b, level) for b in stress], 300003)
        assert not st.any()
        for b, g in zip(stress, got):
            assert g == b, (level, len(b))
    # corrupt streams are flagged, not crashed on
    bad = bytearray(zlib.compress(blobs[-1], 9))
    bad[len(bad) // 2] ^= 0x55
    trunc = zlib.compress(blobs[-1], 9)[:-7]
    _, st = _gpu_inflate(torch_cuda, native, [bytes(bad), trunc, b"\x00\x01\x02\x03\x04\x05\x06"], cap)
    assert (st != 0).all()


# ------------------------------------------------------------------------------------------ end to end

def test_reference_golden_vector(torch_cuda, pkg, golden_cct, golden_digests):
    cfg = pkg.default_config()
    raw = pkg.Decoder(cfg, golden_cct).decode()
    assert hashlib.sha1(raw).hexdigest() == golden_digests["reference_testing"]["decoded_sha1"]
    img = np.frombuffer(raw, dtype=np.uint16).reshape(512, 512)
    enc = pkg.Encoder(cfg, img)
    out = enc.encode()
    assert out == golden_cct
    assert (enc.info["delta"], enc.info["full"], enc.n_jumps, enc.n_difficult) == (248215, 13929, 96, 554)


def test_small_goldens(torch_cuda, pkg, golden_small):
    cfg = pkg.default_config()
    names = sorted(k[:-4] for k in golden_small if k.endswith("__in"))
    for name in names:
        img = golden_small[name + "__in"]
        ref = golden_small[name + "__out"].tobytes()
        assert pkg.Encoder(cfg, img).encode() == ref, name
        if bool(golden_small[name + "__domain"]):
            out = np.frombuffer(pkg.Decoder(cfg, ref).decode(), dtype=np.uint16).reshape(img.shape)
            assert (out == img).all(), name


def test_flag_variants(torch_cuda, pkg, golden_small):
    for name in ("walk_32x64", "synth_ct_64"):
        img = golden_small[name + "__in"]
        for fr in (0, 1):
            for sg in (0, 1):
                for df in (0, 1):
                    cfg = pkg.default_config()
                    cfg["encoder"]["transforms"]["fractal"] = bool(fr)
                    cfg["encoder"]["transforms"]["segmentation"] = bool(sg)
                    cfg["encoder"]["deflate_compression"] = bool(df)
                    ref = golden_small[f"{name}__flags{fr}{sg}{df}"].tobytes()
                    assert pkg.Encoder(cfg, img).encode() == ref, (name, fr, sg, df)
                    out = np.frombuffer(pkg.Decoder(cfg, ref).decode(), dtype=np.uint16).reshape(img.shape)
                    assert (out == img).all(), (name, fr, sg, df)


def test_full_size_digests(torch_cuda, pkg, synth, golden_digests):
    cfg = pkg.default_config()
    for seed, d in golden_digests["synth_ct_512"].items():
        img = synth.synth_ct(int(seed))
        out = pkg.Encoder(cfg, img).encode()
        assert len(out) == d["len"] and hashlib.sha256(out).hexdigest() == d["sha256"], seed
        assert pkg.Decoder(cfg, out).decode() == img.tobytes()


def test_radiograph_2048(torch_cuda, pkg, synth, oracle, golden_digests):
    d = golden_digests["synth_radiograph_2048_seed0"]
    img = synth.synth_radiograph(0, 2048)
    out = pkg.Encoder(pkg.default_config(), img).encode()
    assert len(out) == d["len"] and hashlib.sha256(out).hexdigest() == d["sha256"]
    assert pkg.Decoder(pkg.default_config(), out).decode() == img.tobytes()
    # a batch of two more, device-resident, against the oracle
    imgs = np.stack([synth.synth_radiograph(s, 2048) for s in (1, 2)])
    arena, offsets, status = pkg.encode_batch(torch_cuda.from_numpy(imgs.view(np.int16)).cuda())
    arena, offsets = arena.cpu().numpy(), offsets.cpu().numpy()
    for i in range(2):
        assert arena[offsets[i]:offsets[i + 1]].tobytes() == oracle.encode(imgs[i])


def test_batch_device_resident(torch_cuda, pkg, synth, oracle):
    B = 48
    imgs = synth.synth_ct_batch(range(100, 100 + B), workers=8)
    d_imgs = torch_cuda.from_numpy(imgs.view(np.int16)).cuda()
    arena, offsets, status, stats = pkg.encode_batch(d_imgs, return_stats=True)
    assert not status.cpu().numpy().any()
    h_arena, h_off = arena.cpu().numpy(), offsets.cpu().numpy()
    want_arena, want_off = oracle.encode_batch(imgs, nthreads=8)
    assert (h_off == want_off).all()
    assert h_arena[:h_off[-1]].tobytes() == want_arena.tobytes()
    px, dst = pkg.decode_batch(arena, offsets, (512, 512))
    assert not dst.cpu().numpy().any()
    assert torch_cuda.equal(px.view(torch_cuda.int16), d_imgs)
    # decode of oracle-written files, host buffers in and out
    px2, st2 = pkg.decode_batch(want_arena, want_off, (512, 512))
    assert not st2.any() and (px2 == imgs).all()


def test_multi_pass_small_workspace(torch_cuda, native, pkg, synth, oracle):
    """A workspace that holds only 3 images forces several passes over the same stream."""
    L = native.lib()
    B, n = 10, 128
    imgs = np.stack([synth.synth_ct(s, n) for s in range(B)])
    d_px = torch_cuda.from_numpy(imgs.view(np.int16)).cuda()
    ws = torch_cuda.empty(int(L.cct_encode_workspace_bytes(3, n, n, 7)), dtype=torch_cuda.uint8, device="cuda")
    cap = B * int(L.cct_file_bound(n, n))
    arena = torch_cuda.empty(cap, dtype=torch_cuda.uint8, device="cuda")
    off = torch_cuda.empty(B + 1, dtype=torch_cuda.int64, device="cuda")
    st = torch_cuda.empty(B, dtype=torch_cuda.int32, device="cuda")
    native.check(L.cct_encode_batch(d_px.data_ptr(), B, n, n, 7, arena.data_ptr(), cap, off.data_ptr(), st.data_ptr(),
                                    None, ws.data_ptr(), ws.numel(), torch_cuda.cuda.current_stream().cuda_stream))
    torch_cuda.cuda.synchronize()
    h_arena, h_off = arena.cpu().numpy(), off.cpu().numpy()
    for i in range(B):
        assert h_arena[h_off[i]:h_off[i + 1]].tobytes() == oracle.encode(imgs[i]), i
    dws = torch_cuda.empty(int(L.cct_decode_workspace_bytes(4, n, n)), dtype=torch_cuda.uint8, device="cuda")
    px = torch_cuda.empty((B, n, n), dtype=torch_cuda.int16, device="cuda")
    native.check(L.cct_decode_batch(arena.data_ptr(), off.data_ptr(), B, n, n, px.data_ptr(), st.data_ptr(),
                                    dws.data_ptr(), dws.numel(), torch_cuda.cuda.current_stream().cuda_stream))
    torch_cuda.cuda.synchronize()
    assert not st.cpu().numpy().any() and torch_cuda.equal(px, d_px)
    # too small for even one image -> error code, not a crash
    rc = L.cct_encode_batch(d_px.data_ptr(), B, n, n, 7, arena.data_ptr(), cap, off.data_ptr(), st.data_ptr(), None,
                            ws.data_ptr(), 1024, None)
    assert rc == native.E_WORKSPACE


def test_errors_and_domain(torch_cuda, pkg, oracle, native):
    cfg = pkg.default_config()
    with pytest.raises(ValueError):
        pkg.Encoder(cfg, np.zeros((6, 6), dtype=np.uint16)).encode()  # N % 16 (core.py:245)
    with pytest.raises(ValueError):
        pkg.Decoder(cfg, b"lite" + bytes(40)).decode()  # bad magic (core.py:388-389)
    bad = dict(cfg, encoder=dict(cfg["encoder"], transforms=dict(cfg["encoder"]["transforms"], delta=False)))
    with pytest.raises(NotImplementedError):
        pkg.Encoder(bad, np.zeros((16, 16), dtype=np.uint16)).encode()
    zip_ = dict(cfg, encoder=dict(cfg["encoder"], transforms=dict(cfg["encoder"]["transforms"], zipper=True)))
    with pytest.raises(NotImplementedError):
        pkg.Encoder(zip_, np.zeros((16, 16), dtype=np.uint16)).encode()
    # out of the lossless domain: encoder bytes still exact, status flags it, decoder raises like the reference
    img = np.full((32, 32), 3000, dtype=np.uint16)
    enc = pkg.Encoder(cfg, img)
    out = enc.encode()
    assert out == oracle.encode(img) and not enc.domain_ok
    with pytest.raises(OverflowError):
        oracle.decode(out)
    with pytest.raises(OverflowError):
        pkg.Decoder(cfg, out).decode()
    # corrupt zlib stream -> zlib.error like the reference
    good = bytearray(pkg.Encoder(cfg, np.zeros((32, 32), dtype=np.uint16)).encode())
    good[-2] ^= 0xFF
    with pytest.raises(zlib.error):
        pkg.Decoder(cfg, bytes(good)).decode()
    # arena too small -> error, not an overrun
    L = native.lib()
    arena = np.zeros(64, dtype=np.uint8); off = np.zeros(2, dtype=np.int64); st = np.zeros(1, dtype=np.uint32)
    rc = L.cct_encode_host(img.ctypes.data, 1, 32, 32, 7, arena.ctypes.data, 20, off.ctypes.data, st.ctypes.data, None, -1)
    assert rc == native.E_CAPACITY and (st[0] & native.ST_CAPACITY)


def _tagged_run_image(seed, rows, cols, order):
    """Pixels whose op stream holds long runs of bytes that LOOK like FULL tags: a ramp of +229 per op is the byte
    pair E0 E5, so k steps are 2k tagged bytes in a row (op starts alternate inside them), with run lengths and
    phases that put them across 32-byte steps, 128-byte groups and the 1 KB chunks of k_dec_ops_scan/emit."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    vals = np.empty(n, dtype=np.int64)
    v, k = 0, 0
    while k < n:
        up = int(rng.integers(1, 270))
        for _ in range(up):
            if k >= n or v + 229 > 65535:
                break
            v += 229; vals[k] = v; k += 1
        for _ in range(int(rng.integers(0, 40))):  # plain DELTA ops in between (one byte each): shifts the phase
            if k >= n:
                break
            v = max(0, min(65535, v + int(rng.integers(-60, 61)))); vals[k] = v; k += 1
        while v > 3000 and k < n:  # back down in FULL steps whose low byte is no tag
            v -= 2000; vals[k] = v; k += 1
    img = np.empty(n, dtype=np.uint16)
    img[order] = vals.astype(np.uint16)
    return img.reshape(rows, cols)


def _endless_tagged_image(seed, rows, cols, order):
    """An op stream that is tagged bytes almost throughout: +229 per op (bytes E0 E5) on the way up, -283 (EE E5) on
    the way down, a single DELTA op every few thousand bytes to flip the phase.  Runs of tagged bytes far longer than
    the look-back of k_dec_ops_scan: those chunks are summarised for both entry states."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    vals = np.empty(n, dtype=np.int64)
    v, k, up = 300, 0, True
    next_flip = int(rng.integers(1500, 4000))
    while k < n:
        if k == next_flip:
            v += 1 if v < 60000 else -1
            next_flip += int(rng.integers(1500, 4000))
        elif up:
            if v + 229 > 65535: up = False; continue
            v += 229
        else:
            if v - 283 < 0: up = True; continue
            v -= 283
        vals[k] = v; k += 1
    img = np.empty(n, dtype=np.uint16)
    img[order] = vals.astype(np.uint16)
    return img.reshape(rows, cols)


@pytest.mark.parametrize("shape,flags", [((128, 128), 5), ((128, 128), 4), ((256, 256), 5)])
def test_op_stream_tagged_beyond_look_back(torch_cuda, pkg, oracle, shape, flags):
    rows, cols = shape
    fractal, deflate = bool(flags & 1), bool(flags & 4)
    cfg = pkg.default_config()
    tr = cfg["encoder"]["transforms"]
    tr["fractal"], tr["segmentation"] = fractal, False
    cfg["encoder"]["deflate_compression"] = deflate
    order = oracle.curve(rows, cols).astype(np.int64) if fractal else np.arange(rows * cols)
    for seed in range(2):
        img = _endless_tagged_image(seed, rows, cols, order)
        want = oracle.encode(img, fractal=fractal, seg=False, deflate=deflate)
        payload = zlib.decompress(want[13:]) if deflate else want[13:]
        tagged = (np.frombuffer(payload, dtype=np.uint8) & 0xF0) == 0xE0
        longest = max(len(r) for r in "".join("1" if t else "0" for t in tagged).split("0"))
        assert longest > 2048 + 1024, longest
        assert pkg.Encoder(cfg, img).encode() == want
        assert np.array_equal(oracle.decode(want), img)
        assert pkg.Decoder(cfg, want).decode() == img.tobytes()


@pytest.mark.parametrize("shape,flags", [((128, 128), 7), ((128, 128), 4), ((64, 256), 5), ((512, 512), 6)])
def test_op_stream_tagged_runs(torch_cuda, pkg, oracle, shape, flags):
    """Decode of op streams with long FULL-tag look-alike runs (SURVEY A.5's 2-state automaton at its worst)."""
    rows, cols = shape
    fractal, seg, deflate = bool(flags & 1), bool(flags & 2), bool(flags & 4)
    cfg = pkg.default_config()
    tr = cfg["encoder"]["transforms"]
    tr["fractal"], tr["segmentation"] = fractal, seg
    cfg["encoder"]["deflate_compression"] = deflate
    order = oracle.curve(rows, cols).astype(np.int64) if fractal else np.arange(rows * cols)
    for seed in range(3):
        img = _tagged_run_image(seed, rows, cols, order)
        want = oracle.encode(img, fractal=fractal, seg=seg, deflate=deflate)
        if not seg:  # (meshing reorders the blocks, and with them the runs)
            payload = zlib.decompress(want[13:]) if deflate else want[13:]
            tagged = (np.frombuffer(payload, dtype=np.uint8) & 0xF0) == 0xE0
            longest = max(len(r) for r in "".join("1" if t else "0" for t in tagged).split("0"))
            assert longest > 400, "the generator should produce tagged runs longer than a warp step"
        assert pkg.Encoder(cfg, img).encode() == want
        try:
            want_px = oracle.decode(want)
        except OverflowError:  # meshing can take a ramp out of the reference's lossless domain
            with pytest.raises(OverflowError):
                pkg.Decoder(cfg, want).decode()
            continue
        assert np.array_equal(want_px, img)
        assert pkg.Decoder(cfg, want).decode() == img.tobytes()


def test_second_device_in_one_process(torch_cuda, pkg, native, synth, oracle):
    """One process driving two GPUs: opt-in shared-memory sizes and SM counts are kept per device, the host entry
    points select the device they are given and hand the caller's current device back."""
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    imgs = np.stack([synth.synth_ct(s, 128) for s in range(6)])
    want = [oracle.encode(im) for im in imgs]
    for dev in (0, 1, 0):
        d = torch.from_numpy(imgs.view(np.int16)).to(f"cuda:{dev}")
        arena, offsets, status = pkg.encode_batch(d)
        px, dst = pkg.decode_batch(arena, offsets, (128, 128))
        torch.cuda.synchronize(dev)
        h, off = arena.cpu().numpy(), offsets.cpu().numpy()
        assert [h[off[i]:off[i + 1]].tobytes() for i in range(6)] == want, dev
        assert torch.equal(px.view(torch.int16), d) and not status.any().item() and not dst.any().item()
    # host-buffer entry points on the device that is not current
    L = native.lib()
    torch.cuda.set_device(0)
    cap = int(L.cct_file_bound(128, 128)) * 6
    arena = np.zeros(cap, np.uint8); off = np.zeros(7, np.int64); st = np.zeros(6, np.uint32)
    native.check(L.cct_encode_host(imgs.ctypes.data, 6, 128, 128, 7, arena.ctypes.data, cap, off.ctypes.data, st.ctypes.data, None, 1))
    assert [arena[off[i]:off[i + 1]].tobytes() for i in range(6)] == want
    back = np.zeros_like(imgs); dst = np.zeros(6, np.uint32)
    native.check(L.cct_decode_host(arena.ctypes.data, off.ctypes.data, 6, 128, 128, back.ctypes.data, dst.ctypes.data, 1))
    assert np.array_equal(back, imgs) and not dst.any()
    assert torch.cuda.current_device() == 0


def test_host_entry_points_lanes_and_passes(torch_cuda, native, synth, oracle, monkeypatch):
    """cct_encode_host / cct_decode_host over several passes on one, two and three lanes (CCT_HOST_LANES), with the
    downloads trailing the pass being queued: bytes equal the oracle's whatever the schedule, and the pixels come back."""
    L = native.lib()
    imgs = np.stack([synth.synth_ct(s, 128) for s in range(23)])
    want = [oracle.encode(im) for im in imgs]
    cap = int(L.cct_file_bound(128, 128)) * 23
    for lanes, sizes in (("1", "5,5,5,8"), ("2", "1,2,4,8,8"), ("3", "3,3,3,3,3,3,3,2"), ("3", "23"), ("3", "11,12"), ("2", "")):
        monkeypatch.setenv("CCT_HOST_LANES", lanes)
        if sizes:
            monkeypatch.setenv("CCT_HOST_SIZES", sizes)
            monkeypatch.setenv("CCT_HOST_SIZES_DEC", sizes)
        else:
            monkeypatch.delenv("CCT_HOST_SIZES", raising=False)
            monkeypatch.delenv("CCT_HOST_SIZES_DEC", raising=False)
        arena = np.zeros(cap, np.uint8); off = np.zeros(24, np.int64); st = np.ones(23, np.uint32)
        native.check(L.cct_encode_host(imgs.ctypes.data, 23, 128, 128, 7, arena.ctypes.data, cap, off.ctypes.data, st.ctypes.data, None, 0))
        assert not st.any()
        assert [arena[off[i]:off[i + 1]].tobytes() for i in range(23)] == want, (lanes, sizes)
        back = np.zeros_like(imgs); dst = np.ones(23, np.uint32)
        native.check(L.cct_decode_host(arena.ctypes.data, off.ctypes.data, 23, 128, 128, back.ctypes.data, dst.ctypes.data, 0))
        assert np.array_equal(back, imgs) and not dst.any(), (lanes, sizes)


def test_config_batch_roundtrip_4096(torch_cuda, pkg, synth, oracle):
    """BASELINE.json configs[1] at full size through size-independent properties: every bytestream length and a
    checksum of checksums equal the oracle's, and decode(encode(x)) == x for all 4096 slices."""
    import os
    B = 4096
    base = synth.synth_ct_batch(range(256), workers=min(32, os.cpu_count() or 8))
    imgs = np.concatenate([base] * (B // 256))  # seeds repeat mod 256: bounded CPU time, full-size batch
    d_imgs = torch_cuda.from_numpy(imgs.view(np.int16)).cuda()
    arena, offsets, status = pkg.encode_batch(d_imgs)
    assert not status.cpu().numpy().any()
    h_off = offsets.cpu().numpy()
    want_arena, want_off = oracle.encode_batch(base, nthreads=min(32, os.cpu_count() or 8))
    lens = np.diff(h_off)
    assert (lens.reshape(B // 256, 256) == np.diff(want_off)[None, :]).all()
    h_arena = arena[:h_off[-1]].cpu().numpy()
    assert h_arena[:want_off[-1]].tobytes() == want_arena.tobytes()
    rep = h_arena.reshape(B // 256, -1)
    assert (rep == rep[0]).all()  # identical inputs -> identical bytes in every replica
    px, dst = pkg.decode_batch(arena, offsets, (512, 512))
    assert not dst.cpu().numpy().any()
    assert torch_cuda.equal(px.view(torch_cuda.int16), d_imgs)


@pytest.fixture(scope="module")
def hostsim():
    """The product's scalar cores compiled for the CPU (tests/hostsim): the expected intermediate results."""
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
    subprocess.run(["make", "-C", here], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(here, "libcct_hostsim.so"))
    L.hs_match_tables.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hs_match_tables.restype = None
    return L


def test_stage_match_tables(torch_cuda, native, hostsim, oracle, synth):
    """The match finder of the DEFLATE stage: hash-sorted chain members (k_hs_*) and the longest_match tables
    (k_bprobe) against the same cores run sequentially on the CPU over a reference sort."""
    torch = torch_cuda
    rng = np.random.default_rng(3)
    blobs = [rng.integers(0, a, size=n, dtype=np.uint8).tobytes() for a in (2, 3, 16, 256) for n in (100, 5000, 40000)]
    blobs += [bytes(9000) + rng.integers(0, 4, size=3000, dtype=np.uint8).tobytes() + bytes(70000) + b"abc" * 500]
    blobs += [b"", b"a", b"abc", b"aaaa", bytes(4097), bytes(4096) + b"ab"]
    blobs += [oracle.encode(synth.synth_ct(s, 256), detail=True)[1]["payload"] for s in (0, 1)]
    # byte runs (k_run_tails): shared continuations, both chain budgets, the distance limit, a run at position 0 and at
    # the end, capped run counts, collisions in the run's bucket, and more records than a run keeps (the general-walk
    # fall-back inside k_run_tails)
    from test_host_logic import _runs_blob
    afters = [[5, 6, 7, 8, 9, 1, 2, 3], [5, 6, 7, 1], [5, 1], [9], list(range(10, 250)), list(range(10, 250)) + [3]]
    blobs += [
        _runs_blob(rng, 60, [0], [258, 259, 260, 300, 515, 516, 517, 1000, 257, 100], afters, [1, 2, 5, 30]),
        _runs_blob(rng, 80, [0, 0, 0, 7, 200], [258, 270, 400, 3, 5, 40, 258 + 255], afters, [1, 3, 10]),
        _runs_blob(rng, 30, [0], [2000, 5000, 4096 + 300, 1024 + 258, 1024 + 257, 900], afters, [1, 2]),
        bytes(300) + _runs_blob(rng, 20, [0], [400], afters, [2]) + bytes(600) + b"\x05" * 255,
        _runs_blob(rng, 12, [0], [9000, 12000, 300], afters, [2, 700]),
        bytes(70000) + b"\x05\x06" + bytes(66000) + b"\x05\x06\x07" + bytes(400) + b"\x05\x06",
        b"".join(bytes(300) + bytes([5, 6, 7]) + bytes([32, 0, 0, 1]) * 50 for _ in range(20)),
        b"".join(bytes(300 if i % 25 == 24 else int(rng.integers(3, 60))) + bytes([5, 6, int(rng.integers(1, 4))]) for i in range(600)),
        b"".join(bytes(300 if i % 10 == 9 else int(rng.integers(100, 280))) + bytes([5, int(rng.integers(1, 3))]) for i in range(300)),
    ]
    L = native.lib()
    B = len(blobs)
    max_len = max(len(b) for b in blobs)
    stride = (max_len + 31) // 16 * 16
    h_in = np.zeros((B, stride), dtype=np.uint8)
    for i, b in enumerate(blobs):
        h_in[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    h_in[:, :] |= 0  # (rows are zero past the end here; the kernels must not depend on it: see the garbage pass below)
    for garbage in (False, True):
        if garbage:
            for i, b in enumerate(blobs):
                h_in[i, len(b):] = rng.integers(0, 256, size=stride - len(b), dtype=np.uint8)
        d_in = _dev(torch, h_in)
        d_len = _dev(torch, np.array([len(b) for b in blobs], dtype=np.int32))
        d_ent = torch.zeros((B, stride), dtype=torch.int64, device="cuda")
        d_lo = torch.zeros((B, stride), dtype=torch.int32, device="cuda")
        d_hi = torch.zeros((B, stride), dtype=torch.int32, device="cuda")
        ws = torch.empty(int(L.cct_deflate_workspace_bytes(B, max_len)), dtype=torch.uint8, device="cuda")
        native.check(L.cct_debug_match_tables(d_in.data_ptr(), stride, d_len.data_ptr(), B, max_len, d_ent.data_ptr(),
                                              d_lo.data_ptr(), d_hi.data_ptr(), stride, ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        g_ent = d_ent.cpu().numpy().view(np.uint64)
        g_lo = d_lo.cpu().numpy().view(np.uint32)
        g_hi = d_hi.cpu().numpy().view(np.uint32)
        for i, b in enumerate(blobs):
            n = len(b)
            src = np.frombuffer(b, dtype=np.uint8) if n else np.zeros(1, np.uint8)
            w_ent = np.zeros(max(n, 1), np.uint64)
            w_lo = np.zeros(max(n, 1), np.uint32)
            w_hi = np.zeros(max(n, 1), np.uint32)
            hostsim.hs_match_tables(src.ctypes.data, n, w_ent.ctypes.data, w_lo.ctypes.data, w_hi.ctypes.data)
            m = max(n - 3, 0)
            bad = np.nonzero(g_ent[i, :m] != w_ent[:m])[0]
            assert len(bad) == 0, ("sorted entries", i, n, bad[:4], g_ent[i, bad[:4]], w_ent[bad[:4]])
            bad = np.nonzero(g_lo[i, :n] != w_lo[:n])[0]
            assert len(bad) == 0, ("budget-4096 table", i, n, bad[:4], g_lo[i, bad[:4]], w_lo[bad[:4]])
            flagged = (w_lo[:n] & 0x40000000) != 0
            assert (g_hi[i, :n][flagged] == w_hi[:n][flagged]).all(), ("budget-1024 table", i, n)


def test_evaluate_harness(torch_cuda, pkg, synth, oracle, tmp_path):
    """The ratio study of scripts/evaluate.py on arrays: the CCT column is the reference encoder's size."""
    import importlib
    ev = importlib.import_module(pkg.__name__ + ".evaluate")
    imgs = np.stack([synth.synth_ct(s, 128) for s in range(6)])
    rows = ev.comparison_batch(imgs)
    for i, r in enumerate(rows):
        assert r[ev.RAW] == imgs[i].nbytes
        assert r[ev.CCT] == len(oracle.encode(imgs[i]))
        assert r[ev.ZIP] == len(zlib.compress(imgs[i].tobytes()))
        assert 0 < r[ev.PNG] < r[ev.RAW]
    ev.write_csv(rows, str(tmp_path / "results.csv"))
    lines = open(tmp_path / "results.csv").read().splitlines()
    assert lines[0] == "File,Raw,ZIP,PNG,CCT" and len(lines) == 7
    assert ev.ratios(rows)["CCT"] > ev.ratios(rows)["ZIP"]


def test_two_streams_and_owned_results(torch_cuda, pkg, synth):
    """Calls on two CUDA streams run at the same time without sharing scratch, and a result stays valid after the
    next call (the arena belongs to the caller)."""
    torch = torch_cuda
    base = np.stack([synth.synth_ct(s, 128) for s in range(24)])
    a = torch.from_numpy(np.concatenate([base] * 25).view(np.int16)).cuda()          # 600 images: decode in two halves
    b = torch.from_numpy(np.concatenate([base[::-1]] * 25).view(np.int16).copy()).cuda()
    ref_a = pkg.encode_batch(a)
    ref_b = pkg.encode_batch(b)
    torch.cuda.synchronize()
    want_a = (ref_a[0][:int(ref_a[1][-1])].clone(), ref_a[1].clone())
    want_b = (ref_b[0][:int(ref_b[1][-1])].clone(), ref_b[1].clone())
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            ea = pkg.encode_batch(a)
            da, sa = pkg.decode_batch(ea[0], ea[1], (128, 128))
        with torch.cuda.stream(s2):
            eb = pkg.encode_batch(b)
            db, sb = pkg.decode_batch(eb[0], eb[1], (128, 128))
        torch.cuda.synchronize()
        assert torch.equal(ea[1], want_a[1]) and torch.equal(ea[0][:want_a[0].numel()], want_a[0])
        assert torch.equal(eb[1], want_b[1]) and torch.equal(eb[0][:want_b[0].numel()], want_b[0])
        assert torch.equal(da.view(torch.int16), a) and torch.equal(db.view(torch.int16), b)
        assert not sa.any().item() and not sb.any().item()
    # the first results are still what they were
    assert torch.equal(ref_a[0][:want_a[0].numel()], want_a[0]) and torch.equal(ref_b[0][:want_b[0].numel()], want_b[0])
