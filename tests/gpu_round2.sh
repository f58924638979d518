#!/bin/bash
# One GPU-box pass of round 2 (evidence for profiles/): parity tests, smoke, both bench arms as the driver runs them,
# the default bench line, ncu launch list of resident steps, ncu --set full captures, in-situ layer timelines.
# usage: tests/gpu_round2.sh <tag> [notests]
tag=${1:-r02x}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${tag}_smi.txt 2>&1
nproc >> $o/${tag}_smi.txt
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $o/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $o/${tag}_pytest.log
  tail -4 $o/${tag}_pytest.log
  timeout 600 python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1
  echo "smoke exit $?"; tail -2 $o/${tag}_smoke.log
fi
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $o/${tag}_bench_reference_s20.json 2> $o/${tag}_bench_reference_s20.err
echo "reference arm exit $?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $o/${tag}_bench_s20.json 2> $o/${tag}_bench_s20.err
echo "bench (driver flags) exit $?"; tail -2 $o/${tag}_bench_s20.err
timeout 900 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
echo "bench (default) exit $?"; tail -2 $o/${tag}_bench.err
python - $o/${tag}_bench_s20.json $o/${tag}_bench.json $o/${tag}_bench_reference_s20.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        e = d.get("e2e", {})
        print(f"{f}: value {d['value']:.2f} ms/step {d['ms_per_step']:.3f} e2e {e.get('value', 0):.2f}",
              "sync", e.get("sync_scheduler", {}).get("value"), "gemm", d.get("roofline", {}).get("frac"),
              "attn", d.get("roofline_attention", {}).get("frac"), d.get("roofline_attention", {}).get("avg_launch_us"),
              "ttfa1", d.get("ttfa_single_ms", {}).get("p50"), "join", d.get("ttfa_join_ms", {}).get("p50"),
              "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $o/${tag}_launches.csv python bench.py --steps 14 --warmup 3 --no-cpu --ttfa-joins 0 --profile-steps 7 > $o/${tag}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
for kv in 200 500 900; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn -c 3 -o $o/${tag}_attn_kv${kv}_full -f \
    python tests/prof_attn.py $kv > $o/${tag}_ncu_attn_kv$kv.log 2>&1
  echo "ncu attn kv $kv exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 3 -o $o/${tag}_gemm_full -f \
  python tests/prof_gemm.py > $o/${tag}_ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 300 python tests/prof_mimi.py 64 10 > $o/${tag}_mimi_timing.txt 2>&1; cat $o/${tag}_mimi_timing.txt
timeout 300 python tests/prof_qwen3_codec.py 16 3 > $o/${tag}_qwen3_codec_timing.txt 2>&1; cat $o/${tag}_qwen3_codec_timing.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:snac_gemm_tf32x3 -s 45 -c 4 -o $o/${tag}_codec_tc_full -f \
  python tests/prof_qwen3_codec.py 16 1 > $o/${tag}_ncu_codec.log 2>&1
echo "ncu codec (tcgen05 conv kernel) exit $?"
for b in 64; do timeout 400 python tests/prof_csm.py $b 600 60 2>&1 | grep "^{" > $o/${tag}_csm_b$b.json; done
for b in 16 32; do timeout 300 python tests/prof_qwen3_tts.py $b 50 60 2>&1 | grep "^{" > $o/${tag}_qwen3_tts_b$b.json; done
cat $o/${tag}_csm_b64.json $o/${tag}_qwen3_tts_b16.json | cut -c1-300
for kv in 200 500; do
  timeout 200 python tests/trace_step.py $kv unfused 60 34 > $o/${tag}_trace_kv$kv.txt 2>&1
  head -18 $o/${tag}_trace_kv$kv.txt
done
VB_DECODE_MODE=fused timeout 400 python bench.py --steps 70 --warmup 3 --no-cpu --ttfa-joins 0 > $o/${tag}_bench_fused.json 2> $o/${tag}_bench_fused.err
echo "fused-mode bench exit $?"
ls -la $o | grep ${tag} | awk '{print $5, $9}'
