#!/bin/bash
# Round-2 ncu capture of selected kernels (run on the GPU box from the repository root through gpurun, one GPU).
# Usage: profiles/capture_ncu_kernels.sh <tag> <kernel regex>   -> gpurun_out/<tag>_raw.csv, <tag>_src_<kernel>.csv
set -e
tag=${1:-cap}
rx=${2:-k_bprobe}
A="--batch 1184 --steps 1 --warmup 1 --no-e2e --cpu-sample 32 --profile-all"   # 1184 slices = 8 streams per SM
mkdir -p gpurun_out
python bench.py $A > gpurun_out/${tag}_plain.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"$rx" --launch-skip-before-match 0 -c 12 -f \
    -o gpurun_out/${tag}_prof python bench.py $A > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
for k in $(echo "$rx" | tr '|' ' '); do
  ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv --kernel-name regex:$k --print-source cuda,sass \
      > gpurun_out/${tag}_src_$k.csv 2>/dev/null || true
done
rm -f gpurun_out/${tag}_prof.ncu-rep
