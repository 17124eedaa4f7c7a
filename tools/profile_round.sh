#!/bin/bash
# Run on the GPU box (gpurun): plain bench first, then the ncu launch list and one full capture of the two
# dominant kernels.  Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.txt here.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"rtj_scan_sync_kernel|rtj_scan_chunk_kernel|rtj_idct_kernel|rtj_idct_hard|rtj_resolve" \
    -s 18 -c 6 -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -c 400 gpurun_out/bench_${TAG}.json
