#!/bin/bash
# GEMM bottleneck experiment: 1 = no TMA, 2 = no MMA, 4 = no epilogue (timing only; results are wrong)
for dbg in 0 1 2 4 3 6 5 7; do
  echo -n "ZS_GEMM_DEBUG=$dbg "
  ZS_GEMM_DEBUG=$dbg python bench.py --steps 5 --segments 256 --micro-batch 256 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print('gemm_ms', round(r['kernel_ms_per_step'],2))
except Exception as e: print('failed', e)"
done
