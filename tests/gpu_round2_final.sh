#!/bin/bash
# Final GPU-box pass of round 2 (evidence for profiles/): the whole parity suite, smoke, both bench arms as the driver runs
# them, the default bench line, probes of the prefill kernel / LM variants / speech tokenizer, ncu launch list of resident
# steps and one --set full capture of the tiled prefill kernel.  usage: tests/gpu_round2_final.sh <tag>
tag=${1:-r02f}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
tail -3 $o/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1
echo "smoke exit $?"; grep "^smoke" $o/${tag}_smoke.log | cut -c1-400
timeout 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $o/${tag}_bench_reference_s20.json 2> $o/${tag}_bench_reference_s20.err
echo "reference arm exit $?"
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $o/${tag}_bench_s20.json 2> $o/${tag}_bench_s20.err
echo "bench (driver flags) exit $?"; tail -2 $o/${tag}_bench_s20.err
timeout 600 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
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
              "burst", d.get("ttfa_burst_ms", {}).get("p50"), "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
timeout 100 python tests/prof_prefill_attn.py > $o/${tag}_prefill_attn.jsonl 2> $o/${tag}_prefill_attn.err; echo "prefill probe exit $?"
timeout 100 python tests/prof_attn_time.py 200 500 900 > $o/${tag}_attn_time.jsonl 2>&1; cat $o/${tag}_attn_time.jsonl
(timeout 200 python tests/prof_lm_variants.py cosyvoice2 1 300 200; timeout 300 python tests/prof_lm_variants.py glm 8 435 77) 2> $o/${tag}_lm_variants.err | grep "^{" > $o/${tag}_lm_variants.jsonl; echo "lm variants exit $?"; cut -c1-330 $o/${tag}_lm_variants.jsonl
timeout 60 python tests/prof_glm_encoder.py > $o/${tag}_glm_encoder.json 2>&1; cat $o/${tag}_glm_encoder.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_prefill_attn -s 10 -c 2 -o $o/${tag}_prefill_attn_full -f \
  python tests/prof_prefill_attn.py > $o/${tag}_ncu_prefill.log 2>&1
echo "ncu prefill exit $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $o/${tag}_launches.csv python bench.py --steps 14 --warmup 3 --no-cpu --ttfa-joins 0 --profile-steps 7 > $o/${tag}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  -k regex:gemm_bf16 --log-file $o/${tag}_gemm_dram.csv python bench.py --steps 5 --warmup 3 --no-cpu --ttfa-joins 0 --profile-steps 1 > $o/${tag}_ncu_gemm_dram.log 2>&1
echo "ncu gemm dram exit $?"
ls -la $o | grep ${tag} | awk '{print $5, $9}'
