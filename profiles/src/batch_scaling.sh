#!/bin/bash
# Device-resident throughput against the batch size: per-kernel times scaled to 4096 slices (what a pass of the
# host path costs beyond its share).  Run on the GPU box from the repository root.
for b in 427 854 1708 4096; do
  python bench.py --batch $b --steps 5 --warmup 3 --no-e2e --cpu-sample 16 > /tmp/b.json 2>/dev/null
  python - $b <<'P'
import json, sys
b = int(sys.argv[1])
a = json.loads(open("/tmp/b.json").read().strip().splitlines()[-1]); k = a["kernels_ms_warmup_step"]
names = ("k_bprobe", "k_hs_bins", "k_hs_split", "k_block_build", "k_pack", "k_symscan", "k_mesh", "k_emit", "k_walk_stitch", "k_walk_spec", "k_next", "k_run_tails", "k_gather_stats", "k_hs_scan")
print(b, "enc %.0f dec %.0f MPix/s" % (a["encode_mpix_s"], a["decode_mpix_s"]), "enc ms per 4096: %.1f" % (4096 * 262144 / a["encode_mpix_s"] / 1e3),
      {n: round(k[n] * 4096 / b, 2) for n in names if n in k})
P
done
