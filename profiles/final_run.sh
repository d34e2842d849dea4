#!/bin/bash
# The measurements kept under profiles/ for a round (run on the GPU box from the repository root, one GPU):
# GPU tests, the bench line, the reference arm, the radiograph and decode-sweep lines, then the ncu captures.
# Usage: profiles/final_run.sh <round tag, e.g. r02>
round=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > profiles/${round}_pytest_gpu.log 2>&1; tail -2 profiles/${round}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > profiles/${round}_bench_4096.json 2> gpurun_out/${round}_bench.err; tail -c 400 profiles/${round}_bench_4096.json
python bench.py --impl reference --steps 5 --warmup 1 > profiles/${round}_bench_reference_arm.json 2>> gpurun_out/${round}_bench.err
python bench.py --workload radiograph --size 2048 --batch 64 --steps 5 --warmup 3 --cpu-sample 4 > profiles/${round}_bench_radiograph2048_b64.json 2>> gpurun_out/${round}_bench.err
python bench.py --decode-sweep --steps 5 --warmup 3 > profiles/${round}_bench_decode_sweep.json 2>> gpurun_out/${round}_bench.err
bash profiles/capture_ncu.sh ${round} fin
cp profiles/${round}_* gpurun_out/
