#!/bin/bash
# One GPU-box pass of round 2: parity tests, prefetcher ablation, traces.   usage: tests/gpu_round2.sh <tag> [quick]
tag=${1:-r02x}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${tag}_smi.txt 2>&1
# the new / changed tests first, then the rest
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_csm.py tests/test_gpu_qwen3_tts.py \
  tests/test_gpu_e2e.py tests/test_gpu_reference_dropin.py tests/test_gpu_vocoder_graph.py tests/test_gpu_true_dims.py \
  -q --timeout 600 -p no:cacheprovider -s > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
grep -E "passed|failed|FAILED|true-dims|reference adapter|csm |qwen3" $o/${tag}_pytest.log | tail -20
run_bench() {   # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 70 --warmup 3 --no-cpu --ttfa-joins 0 > $o/${tag}_bench_$name.json 2> $o/${tag}_bench_$name.err
  python - "$o/${tag}_bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:>14}: value {d['value']:.1f} ms/step {d['ms_per_step']:.3f} e2e {d['e2e']['value']:.1f} "
          f"sync {d['e2e']['sync_scheduler']['value']:.1f} attn {d['roofline_attention']['frac']:.3f} "
          f"({d['roofline_attention']['avg_launch_us']:.1f} us) gemm {d['roofline']['frac']:.3f} ttfa1 {d['ttfa_single_ms']['p50']:.1f}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run_bench pf_all VB_L2_PREFETCH=1
run_bench pf_off VB_L2_PREFETCH=0
run_bench pf_w_only VB_L2_PREFETCH_KV=0
run_bench pf_win32 VB_L2_WINDOW_MB=32
run_bench pf_win96 VB_L2_WINDOW_MB=96
run_bench pf_attn111 VB_ATTN_SMEM_KB=111
tail -3 $o/${tag}_bench_pf_all.err
timeout 200 python tests/trace_step.py 500 unfused 60 34 > $o/${tag}_trace_kv500.txt 2>&1
VB_L2_PREFETCH=0 timeout 200 python tests/trace_step.py 500 unfused 60 34 > $o/${tag}_trace_kv500_nopf.txt 2>&1
head -40 $o/${tag}_trace_kv500.txt
