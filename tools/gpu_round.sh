#!/bin/bash
# One gpurun call: GPU parity tests, the bench (both arms), the ncu launch list and --set full captures.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh <tag> [quick]
tag=${1:-r1}
quick=${2:-}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/gpu_$tag.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_$tag.log
tail -3 $out/pytest_gpu_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
python tools/bench_summary.py $out/bench_$tag.json
if [ -z "$quick" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
  # launch list of the same command (cold-cache, serialised: shares only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches_$tag.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/launches_$tag.log 2>&1
  # one DiT block (qkv, attention, proj, ln, fc1, fc2, ln ...) with the full metric set
  timeout 900 ncu --set full --clock-control none --import-source on -s 300 -c 10 -o $out/prof_dit_block_$tag -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_block_$tag.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dit_step -s 1 -c 1 -o $out/prof_dit_step_$tag -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_step_$tag.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gin_ -s 40 -c 14 -o $out/prof_gin_$tag -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_gin_$tag.log 2>&1
fi
ls -la $out | tail -20
