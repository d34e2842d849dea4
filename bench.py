#!/usr/bin/env python
"""bench.py — CompaCT encode+decode throughput on B200 (BASELINE.json metric), one process per GPU.

A step = one pass of the hot path over one batch: encode B slices, then decode the B bytestreams.
  value        MPix/s through encode -> decode with pixels and bytestreams resident in HBM (CUDA events)
  e2e          same metric through the host-buffer C ABI (pinned host pixels in, host pixels out; H2D/D2H inside)
  roofline     the dominant kernel's algorithmic bytes per launch / its CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline the oracle port (C restatement + libz) on the host cores, on a bounded sample
`--impl reference` times the reference's CPU implementation: the reference is pure Python and is not present on
the GPU box, so this arm runs the oracle port (kind "port") on all host threads.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "2023-compact-image-compression_b200"

METRIC = "encode+decode MPix/s (12-bit 512x512 CT batch)"  # BASELINE.json metric, one number: pixels / (t_encode + t_decode)


WORKLOAD = "ct"  # "ct": synth_ct slices (configs[1]); "radiograph": 16-bit synth_radiograph (configs[3])


def _gen_seeds(seeds, n):
    synth = importlib.import_module(PKG + ".synth")
    if WORKLOAD == "radiograph":
        return np.stack([synth.synth_radiograph(int(s), n) for s in seeds]) if len(seeds) else np.empty((0, n, n), np.uint16)
    return synth.synth_ct_batch(seeds, n)


def gen_batch(seeds, n, workers):
    """Synthetic slices for `seeds`, generated in worker processes (called before CUDA is initialised)."""
    seeds = list(seeds)
    if workers <= 1 or len(seeds) < 8:
        return _gen_seeds(seeds, n)
    from concurrent.futures import ProcessPoolExecutor
    chunks = [seeds[i::workers] for i in range(workers)]
    out = np.empty((len(seeds), n, n), dtype=np.uint16)
    with ProcessPoolExecutor(workers) as ex:
        for w, arr in enumerate(ex.map(_gen_chunk, [(c, n, WORKLOAD) for c in chunks])):
            out[w::workers] = arr
    return out


def _gen_chunk(args):
    seeds, n, workload = args
    global WORKLOAD
    WORKLOAD = workload
    return _gen_seeds(seeds, n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_run(imgs, nthreads):
    """One encode+decode pass of the oracle port over imgs on nthreads host threads -> (t_enc, t_dec)."""
    from oracle import oracle
    t0 = time.perf_counter()
    arena, off = oracle.encode_batch(imgs, nthreads)
    t1 = time.perf_counter()
    out = oracle.decode_batch(arena, off, imgs.shape[1:], nthreads)
    t2 = time.perf_counter()
    assert (out == imgs).all()
    return t1 - t0, t2 - t1


def stage_of(kernel):
    if kernel.startswith("m_zero"):
        return "front"
    if kernel in ("k_gather_stats", "k_mesh", "k_emit"):
        return "front"        # reads 2N, writes P
    if kernel.startswith(("k_inflate", "k_lz_expand", "k_adler_check")):
        return "inflate"      # reads F, writes P (K6)
    if kernel.startswith("k_dec_"):
        return "back"         # reads P, writes 2N (K7-K8)
    return "deflate"          # reads P, writes F (K5)


def distinct_slices(B, workers, sweep=False):
    """Distinct synthetic slices generated per GPU (the batch tiles them when the host is too small to generate all)."""
    distinct = min(B, 512) if sweep else B
    if B * 0.04 / workers > 90:  # keep generation within ~1.5 min on small hosts: tile a smaller distinct set
        distinct = max(64, int(90 * workers / 0.04) // 64 * 64)
    return min(distinct, B)


def workload_config(B, n, world, distinct):
    """The `config` object of the JSON line: the same for both arms (the reference arm times a sample of it)."""
    raw = 2 * B * n * n
    return {"workload": (f"{B} synth_ct {n}x{n} 12-bit slices per GPU (configs[1]), encode then decode" if WORKLOAD == "ct"
                         else f"{B} synth_radiograph {n}x{n} 16-bit slices per GPU (configs[3]), encode then decode"),
            "slices_per_gpu": B, "distinct_slices_per_gpu": int(distinct), "l2": f"inputs larger than L2 ({raw / 1e9:.2f} GB raw per GPU)",
            "parallelism": f"batch sharded over {world} GPU(s), no collective on the data path"}


def bind_near_gpu(torch, local):
    """Several ranks on one box: keep this rank (and the pinned buffers it is about to allocate) on the CPUs next to
    its GPU, as a launcher with NUMA binding would.  Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(props.uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0".encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1} & os.sched_getaffinity(0)
        if not cpus:
            return "unbound (no CPU affinity reported)"
        os.sched_setaffinity(0, cpus)
        return f"rank bound to the {len(cpus)} CPUs next to its GPU"
    except Exception as e:  # no NVML, no permission ...: run unbound
        return f"unbound ({type(e).__name__})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="slices per GPU (weak scaling)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cpu-sample", type=int, default=512)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-all", action="store_true",
                    help="event-time every kernel inside the timed region (no extra warm-up step; used by profiles/capture_ncu.sh)")
    ap.add_argument("--workload", default="ct", choices=["ct", "radiograph"],
                    help="ct: synth_ct 12-bit slices (configs[1], the headline); radiograph: 16-bit 2048x2048 (configs[3])")
    ap.add_argument("--decode-sweep", action="store_true",
                    help="configs[4]: decode-only MPix/s over batch sizes 1..16384 (one JSON line with a 'sweep' list)")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.workload == "radiograph" and args.size == 512 and args.batch == 4096:
        args.size, args.batch = 2048, 64

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n, B = args.size, args.batch
    n_px = n * n
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        nthreads = cores
        S = min(B, max(64, 4 * nthreads))  # bounded sample: a few seconds per step on any core count
        imgs = gen_batch(range(S), n, min(32, cores))
        for _ in range(args.warmup):
            cpu_port_run(imgs[:max(nthreads, 8)], nthreads)
        te = td = 0.0
        for _ in range(args.steps):
            a, b = cpu_port_run(imgs, nthreads)
            te += a; td += b
        val = S * n_px * args.steps / (te + td) / 1e6
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "MPix/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": (te + td) / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
                "config": workload_config(B, n, world, distinct_slices(B, max(1, min(32, cores // max(world, 1))))),
                "encode_mpix_s": S * n_px * args.steps / te / 1e6, "decode_mpix_s": S * n_px * args.steps / td / 1e6,
                "cpu_baseline": {"value": val, "unit": "MPix/s", "cores": nthreads, "kind": "port",
                                 "sample": f"{S} synth_ct slices x {args.steps} steps, oracle C port + libz, {nthreads} threads"},
                "e2e": {"value": val, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ---- our arm -------------------------------------------------------------------------------------------
    workers = max(1, min(32, cores // max(world, 1)))
    distinct = distinct_slices(B, workers, args.decode_sweep)
    t0 = time.time()
    base = gen_batch([(rank * B + i) % (1 << 31) for i in range(min(distinct, B))], n, workers)
    imgs = base if len(base) == B else np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B]
    gen_s = time.time() - t0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_near_gpu(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    native = pkg._native
    L = native.lib()

    d_imgs = torch.from_numpy(imgs.view(np.int16)).to(dev)

    def step():
        arena, offsets, status, stats = pkg.encode_batch(d_imgs, return_stats=True)
        px, dstatus = pkg.decode_batch(arena, offsets, (n, n))
        return arena, offsets, status, stats, px, dstatus

    for _ in range(max(args.warmup, 1)):
        arena, offsets, status, stats, px, dstatus = step()
    torch.cuda.synchronize()
    # parity guard inside the bench: round trip on the whole batch + oracle bytes on a few slices
    assert not status.any().item() and not dstatus.any().item()
    assert torch.equal(px.view(torch.int16), d_imgs), "decode(encode(x)) != x"
    h_off = offsets.cpu().numpy()
    F_total = int(h_off[-1])
    st = stats.cpu().numpy().astype(np.int64)
    P_total = int((st[:, 0] + 2 * st[:, 1] + st[:, 2] + 1).sum())
    parity = None
    if rank == 0:
        from oracle import oracle
        h_arena = arena[:h_off[4]].cpu().numpy()
        for i in range(4):
            assert h_arena[h_off[i]:h_off[i + 1]].tobytes() == oracle.encode(imgs[i]), "bytestream differs from oracle"
        parity = "4 bytestreams == oracle; round trip exact on the whole batch"

    if args.decode_sweep:  # configs[4]: decode-only throughput vs batch size (seeds repeat beyond the distinct set)
        sweep = []
        for Bs in (1, 4, 16, 64, 256, 1024, 4096, 16384):
            reps = (Bs + B - 1) // B
            d_sub = d_imgs.repeat(reps, 1, 1)[:Bs].contiguous() if Bs > B else d_imgs[:Bs]
            try:
                a_s, o_s, _ = pkg.encode_batch(d_sub)
                for _ in range(2):
                    px_s, dst_s = pkg.decode_batch(a_s, o_s, (n, n))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    px_s, dst_s = pkg.decode_batch(a_s, o_s, (n, n))
                e1.record()
                torch.cuda.synchronize()
                assert torch.equal(px_s.view(torch.int16), d_sub) and not dst_s.any().item()
                sweep.append({"batch": Bs, "decode_mpix_s": Bs * n_px * args.steps / (e0.elapsed_time(e1) / 1e3) / 1e6})
            except torch.OutOfMemoryError:
                sweep.append({"batch": Bs, "decode_mpix_s": None, "note": "out of memory"})
            del d_sub
        if rank == 0:
            print(json.dumps({"metric": "decode-only MPix/s vs batch size (configs[4])", "unit": "MPix/s", "n_gpus": world,
                              "steps": args.steps, "data": "synthetic", "config": {"workload": f"synth_{WORKLOAD} {n}x{n}"},
                              "sweep": sweep}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # per-kernel CUDA events: every kernel during one extra warm-up step, only the largest few inside the timed region
    # (two event records around each of the ~115 launches of a step cost about 2 % of it)
    L.cct_profile_select(None)
    prof_all = {}
    if not args.profile_all:
        L.cct_profile_enable(1)
        step()
        L.cct_profile_enable(0)
        prof_all = native.profile_read()
        top_names = [k for k, _ in sorted(((k, v) for k, v in prof_all.items() if k.startswith("k_")), key=lambda kv: -kv[1][1])[:6]]
        L.cct_profile_select(",".join(top_names).encode())
    sampler = ClockSampler(local)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    L.cct_launch_count(1)
    L.cct_profile_enable(1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    for k in range(args.steps):
        ev[k][0].record()
        arena, offsets, status, stats = pkg.encode_batch(d_imgs, return_stats=True)
        ev[k][1].record()
        px, dstatus = pkg.decode_batch(arena, offsets, (n, n))
        ev[k][2].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    L.cct_profile_enable(0)
    L.cct_profile_select(None)
    launches = int(L.cct_launch_count(0))  # kernels + the library's memsets (m_zero_*)
    prof = native.profile_read()
    t_enc = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)) / 1e3
    t_dec = sum(ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)) / 1e3
    t_all = ev[0][0].elapsed_time(ev[-1][2]) / 1e3
    times = torch.tensor([t_all, t_enc, t_dec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.barrier()
    t_all, t_enc, t_dec = (float(x) for x in times.cpu())
    total_px = world * B * n_px * args.steps

    # ---- end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        h_px = torch.from_numpy(imgs.view(np.int16)).pin_memory()
        cap = B * int(L.cct_file_bound(n, n))
        h_arena_t = torch.empty(min(cap, int(F_total * 1.05) + 4096), dtype=torch.uint8).pin_memory()
        h_off_t = torch.zeros(B + 1, dtype=torch.int64).pin_memory()
        h_st = torch.zeros(B, dtype=torch.int32).pin_memory()
        h_out = torch.empty((B, n, n), dtype=torch.int16).pin_memory()

        split = [0.0, 0.0]

        def e2e_step():
            ta = time.perf_counter()
            native.check(L.cct_encode_host(h_px.data_ptr(), B, n, n, 7, h_arena_t.data_ptr(), h_arena_t.numel(),
                                           h_off_t.data_ptr(), h_st.data_ptr(), None, local))
            tb = time.perf_counter()
            native.check(L.cct_decode_host(h_arena_t.data_ptr(), h_off_t.data_ptr(), B, n, n, h_out.data_ptr(),
                                           h_st.data_ptr(), local))
            split[0] += tb - ta; split[1] += time.perf_counter() - tb
        del arena, px  # free torch's cached copies before the library allocates its own context
        pkg.codec.core._DeviceScratch.clear()
        torch.cuda.empty_cache()
        e2e_step()
        assert torch.equal(h_out, h_px)
        # the link the end-to-end number is bounded by: pinned-memory copy rates of this box
        d_tmp = torch.empty_like(d_imgs)
        link = []
        for src, dst in ((h_px, d_tmp), (d_tmp, h_out)):
            torch.cuda.synchronize(); tc = time.perf_counter()
            dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
            link.append(src.numel() * 2 / (time.perf_counter() - tc) / 1e9)
        del d_tmp
        split[0] = split[1] = 0.0
        if world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        t_e2e = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": total_px / float(t_e2e.item()) / 1e6, "unit": "MPix/s",
               "h2d_bytes_per_step": B * n_px * 2 + F_total + (B + 1) * 8,
               "d2h_bytes_per_step": F_total + (B + 1) * 8 + B * 4 + B * n_px * 2 + B * 4,
               "encode_ms_per_step": split[0] / args.steps * 1e3, "decode_ms_per_step": split[1] / args.steps * 1e3,
               "pinned_copy_gbs": {"h2d": link[0], "d2h": link[1]}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    raw = 2 * B * n_px
    stage_bytes = {"front": raw + P_total, "deflate": P_total + F_total, "inflate": F_total + P_total, "back": P_total + raw}
    top = max(((k, v) for k, v in prof.items() if k.startswith("k_")), key=lambda kv: kv[1][1])
    top_name, (top_n, top_ms) = top
    per_launch_s = top_ms / top_n / 1e3
    achieved = stage_bytes[stage_of(top_name)] / per_launch_s / 1e9
    # DRAM bytes per launch of that kernel, from the committed ncu --set full capture (per stream x this batch)
    traffic = None
    tr_files = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json")) \
        if os.path.isdir(os.path.join(ROOT, "profiles")) else []
    if tr_files:
        tr = json.load(open(os.path.join(ROOT, "profiles", tr_files[-1]))).get("kernels", {})
        if top_name in tr:
            traffic = tr[top_name]["dram_bytes_per_stream"] * B * n_px / (512 * 512)
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": stage_bytes[stage_of(top_name)],
                "avg_launch_ms": top_ms / top_n, "share_of_step": top_ms / (t_all * 1e3)}
    path_bytes = (raw + F_total) * args.steps  # 2N + F per image, encode and decode alike
    path = {"encode_gbs": path_bytes / t_enc / 1e9, "decode_gbs": path_bytes / t_dec / 1e9,
            "encode_frac": path_bytes / t_enc / 1e9 / peak, "decode_frac": path_bytes / t_dec / 1e9 / peak}

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample
    S = min(B, args.cpu_sample)
    cpu_port_run(imgs[:min(S, 2 * cores)], cores)
    ce, cd = cpu_port_run(imgs[:S], cores)
    cpu = {"value": S * n_px / (ce + cd) / 1e6, "unit": "MPix/s", "cores": cores, "kind": "port",
           "sample": f"{S} of the batch's slices, one encode+decode pass, oracle C port + libz on {cores} threads",
           "encode_mpix_s": S * n_px / ce / 1e6, "decode_mpix_s": S * n_px / cd / 1e6}

    line = {"metric": METRIC, "value": total_px / t_all / 1e6, "unit": "MPix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": workload_config(B, n, world, len(base)),
            "encode_mpix_s": total_px / t_enc / 1e6, "decode_mpix_s": total_px / t_dec / 1e6,
            "raw_gbs": {"encode": 2 * total_px / t_enc / 1e9, "decode": 2 * total_px / t_dec / 1e9},
            "compression_ratio": raw / F_total, "roofline": roofline, "roofline_path": path, "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity,
            "kernels_ms_per_step": {k: round(v[1] / args.steps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
            "kernels_ms_warmup_step": {k: round(v[1], 3) for k, v in sorted(prof_all.items(), key=lambda kv: -kv[1][1])},
            "gen_seconds": round(gen_s, 1)}
    if numa:
        line["host_binding"] = numa
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
