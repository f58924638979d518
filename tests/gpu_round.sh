#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list of the resident loop, ncu full capture of the attention kernel.
# usage: tests/gpu_round.sh <tag>
tag=${1:-rXX}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${tag}_smi.txt 2>&1
nproc >> $o/${tag}_smi.txt
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $o/${tag}_pytest.log
tail -5 $o/${tag}_pytest.log
timeout 600 python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1
echo "smoke exit $?"; tail -2 $o/${tag}_smoke.log
timeout 900 python bench.py --steps 700 --warmup 7 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
echo "bench exit $?"; tail -3 $o/${tag}_bench.err; cat $o/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $o/${tag}_launches.csv python bench.py --steps 14 --warmup 3 --no-cpu --profile-steps 7 > $o/${tag}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn -c 3 -o $o/${tag}_attn_full -f \
  python tests/prof_attn.py 728 > $o/${tag}_ncu_attn.log 2>&1
echo "ncu attn exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 3 -o $o/${tag}_gemm_full -f \
  python tests/prof_gemm.py > $o/${tag}_ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:snac_gemm_tf32x3 -s 34 -c 4 -o $o/${tag}_snac_full -f \
  python tests/prof_snac.py 32 1 > $o/${tag}_ncu_snac.log 2>&1
echo "ncu snac exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sample_kernel -c 2 -o $o/${tag}_sampler_full -f \
  python tests/prof_sampler.py > $o/${tag}_ncu_sampler.log 2>&1
echo "ncu sampler exit $?"
