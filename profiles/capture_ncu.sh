#!/bin/bash
# How the ncu evidence under profiles/ is captured (run on the GPU box from the repository root, e.g. through gpurun;
# one GPU, never a multi-rank command).  Usage: profiles/capture_ncu.sh <round tag, e.g. r02> [tag for gpurun_out]
#   1. the same bench command without ncu must exit 0 first;
#   2. launch list: per-launch durations of every kernel of the process (cold-cache, serialised: compare SHARES);
#   3. full capture of the timed step only (--launch-skip = the warm-up half of the launches), exported as the raw
#      page; the .ncu-rep itself is dropped (it exceeds what gpurun brings back);
#   4. profiles/summarize_ncu.py turns the two CSVs into <round>_ncu_b1184_launch*.csv, <round>_ncu_full_b1184.txt and
#      <round>_traffic.json (read by bench.py for roofline.traffic).
set -e
round=${1:-r02}
tag=${2:-cap}
A="--batch 1184 --steps 1 --warmup 1 --no-e2e --cpu-sample 32 --profile-all"   # 1184 slices = 8 streams per SM
mkdir -p gpurun_out
python bench.py $A > gpurun_out/${tag}_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py $A > gpurun_out/${tag}_ncu1.log 2>&1
n=$(grep -c '^"[0-9]' gpurun_out/${tag}_launches.csv)
skip=$(( (n + 1) / 2 ))
ncu --set full --clock-control none --import-source on --launch-skip $skip -f -o gpurun_out/${tag}_prof \
    python bench.py $A > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
for k in k_bprobe k_run_tails k_block_build k_hs_bins; do
  ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv --kernel-name regex:$k --print-source cuda,sass \
      > gpurun_out/${tag}_src_$k.csv 2>/dev/null || true
done
rm -f gpurun_out/${tag}_prof.ncu-rep
N="bench.py $A, $(date -u +%F)"
python profiles/summarize_ncu.py launches gpurun_out/${tag}_launches.csv profiles/${round}_ncu_b1184 \
    "ncu --metrics gpu__time_duration.sum --clock-control none; $N"
python profiles/summarize_ncu.py full gpurun_out/${tag}_raw.csv profiles/${round}_ncu_full_b1184.txt profiles/${round}_traffic.json 1184 \
    "ncu --set full --clock-control none --import-source on --launch-skip $skip (the timed step: every kernel of one encode + decode); $N"
cp profiles/${round}_ncu_b1184_launches.csv profiles/${round}_ncu_b1184_launch_shares.csv profiles/${round}_ncu_full_b1184.txt profiles/${round}_traffic.json gpurun_out/
