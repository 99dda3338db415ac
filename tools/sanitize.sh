#!/bin/bash
# compute-sanitizer pass over the kernels with hand-written synchronisation (CTA-pair GEMM, GEMM + LayerNorm pair / cluster kernels with
# their L2 mailboxes and DSMEM exchange, the fused GIN tail, the fused head epilogue, attention).  Usage on the GPU box: bash tools/sanitize.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
SEL='(gemm_matches_torch and 200-1024-320) or (gemm_ln_residual_matches_torch and (300-1024-1024 or 777-512-256 or 600-1024-256)) or dit_denoiser_logits or dit_reverse_steps_teacher_forced or gin_encoder_matches_reference or gin_predictor_matches_reference or (softmax_topk and 5-20000-50)'
BIG='dit_wide_batch_vs_oracle or gin_encoder_baseline_shape or predictor_head_full_width'
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $out/sanitizer_${tool}_$tag.log 2>&1
  echo "$tool rc=$?" >> $out/sanitizer_${tool}_$tag.log
  tail -4 $out/sanitizer_${tool}_$tag.log
done
# the benchmark-shape kernels (CTA-pair GEMM, pair GEMM + LayerNorm, fused GIN tail, fused head) under memcheck and synccheck
for tool in memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest tests/test_gpu_parity_large.py -m gpu -x -q -k "$BIG" > $out/sanitizer_${tool}_big_$tag.log 2>&1
  echo "$tool rc=$?" >> $out/sanitizer_${tool}_big_$tag.log
  tail -4 $out/sanitizer_${tool}_big_$tag.log
done
# the latency regime: CUDA-graph replay, programmatic dependent launches (prefetches ahead of the dependency wait), split-K + row kernel
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph_replay or full_size_logits or degenerate" > $out/sanitizer_${tool}_lat_$tag.log 2>&1
  echo "$tool rc=$?" >> $out/sanitizer_${tool}_lat_$tag.log
  tail -4 $out/sanitizer_${tool}_lat_$tag.log
done
