for dbg in 0 4 1 2; do
  ZS_GEMM_DEBUG=$dbg timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2o_dbg$dbg.json 2> gpurun_out/r2o_dbg$dbg.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2o_dbg$dbg.json')); r=d['roofline']
    print('dbg=$dbg ms/step %.1f gemm/call %.3f gru %.2f clocks %s power %s' % (d['ms_per_step'], r['kernel_ms_per_960_segment_call'], r['gru_ms_per_step']/20, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')))
except Exception as e: print('dbg=$dbg failed', e)
PY
done
