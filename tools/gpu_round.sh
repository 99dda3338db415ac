#!/bin/bash
# One gpurun call: GPU parity tests, the bench (both arms), the ncu launch list and --set full captures.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh <tag> [quick|notest]   (quick: tests + bench only; notest: no pytest)
tag=${1:-r1}
quick=${2:-}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/gpu_$tag.txt 2>&1
rm -f $out/parity.json
if [ "$quick" != "notest" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -s > $out/pytest_gpu_$tag.log 2>&1
  echo "pytest rc=$?" >> $out/pytest_gpu_$tag.log
  tail -3 $out/pytest_gpu_$tag.log
fi
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
python tools/bench_summary.py $out/bench_$tag.json
if [ "$quick" != "quick" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
  # launch list of the same command (cold-cache, serialised: shares only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches_$tag.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency > $out/launches_$tag.log 2>&1
  # one DiT block (qkv, attention, proj, ln, fc1, fc2, ln ...) with the full metric set
  timeout 900 ncu --set full --clock-control none -s 300 -c 8 -o $out/prof_dit_block_$tag -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency > $out/ncu_block_$tag.log 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:dit_step -s 1 -c 1 -o $out/prof_dit_step_$tag -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency > $out/ncu_step_$tag.log 2>&1
  # one GIN encoder layer set (aggregate, statistics GEMM, mlp0 with LayerNorm + GELU epilogue, fused GEMM + layer tail, pooling)
  timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"gin_aggregate|gin_pool|gemm_ln_pair_kernel|EpiRowSq|EpiLnGelu" -s 12 -c 12 -o $out/prof_gin_$tag -f \
      python bench.py --only gin > $out/ncu_gin_$tag.log 2>&1
  # the fused predictor head (pilot GEMM, head GEMM with the top-k epilogue, select kernel)
  timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"EpiHeadTopk|gin_head_" -c 6 -o $out/prof_head_$tag -f \
      python bench.py --only predictor > $out/ncu_head_$tag.log 2>&1
  cp $out/parity.json $out/parity_$tag.json 2>/dev/null
fi
# summarise on the box (the .ncu-rep files are too large to travel: gpurun_out/ is capped at 64 MiB), keep only the text
if [ "$quick" != "quick" ]; then
  LLB_PROFILES_OUT=$out/profiles python tools/make_profiles.py $tag r2 > $out/make_profiles_$tag.log 2>&1
  rm -f $out/prof_*_$tag.ncu-rep
fi
du -sh $out; ls -la $out | tail -20
