#!/bin/bash
# Run the GPU parity tests one group per process so a faulting kernel cannot poison later groups.
# usage: tests/run_gpu_groups.sh [outfile]
out=${1:-gpurun_out/gpu_groups.log}
mkdir -p "$(dirname "$out")"
: > "$out"
for k in rmsnorm rope plan_rows "paged_attention_decode_orpheus" "paged_attention_decode_batch32" \
         "paged_attention_decode_variants" "paged_attention_prefill" gemm_partials gemm_lm_head gemm_gate_up \
         reduce_residual qkv_rope embedding_gather sampler_golden sampler_greedy sampler_stochastic snac_decode; do
  echo "=== $k" >> "$out"
  timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "$k" --timeout 240 -p no:cacheprovider 2>&1 \
    | grep -E "passed|failed|error|Error|assert|rel l2|max err|ulp|FAILED|trap|illegal|CUDA" | head -40 >> "$out"
done
cat "$out"
