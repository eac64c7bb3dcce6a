#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one full capture of the dominant kernel.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt
nproc >> gpurun_out/${tag}_smi.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${tag}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-rows > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_track_level' --launch-skip 12 -c 4 -o gpurun_out/${tag}_track \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-rows >> gpurun_out/${tag}_ncu_bench.log 2>&1
# gpurun brings back at most 64 MiB: the report travels gzip-compressed (gunzip it before `ncu -i`)
gzip -1 -f gpurun_out/${tag}_track.ncu-rep
ls -la gpurun_out
