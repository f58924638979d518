#!/bin/bash
# dev: gate/up tile-height / ring experiments on the per-layer critical path
out=gpurun_out/r02i_gu_tiles.txt; : > $out
run() { echo "=== $*" >> $out; env "$@" timeout 200 python tests/trace_step.py 200 unfused 0 0 >> $out 2>&1; }
run A=default
run VB_GU_HALF=32 VB_GEMM_SMEM_KB_GU=104
run VB_GU_HALF=32 VB_GEMM_SMEM_KB_GU=112
run VB_GU_HALF=64 VB_GEMM_SMEM_KB_GU=160
run VB_GU_HALF=64 VB_GEMM_SMEM_KB_GU=104
tail -100 $out
