# A/B of the GRU recurrence experiment bits inside the bench (continuous load, MB = 224)
for dbg in 0 32 64 96 4 2 1; do
  ZS_GRU_DEBUG=$dbg ZS_GRU_FAST_ACT=1 timeout 100 python bench.py --steps 10 --no-cpu-baseline --segments 224 --micro-batch 224 2>/dev/null | python tools/print_bench.py dbg$dbg
done
for dbg in 8 40 72 104; do ZS_GRU_FAST_ACT=1 ZS_GRU_DEBUG=$dbg timeout 60 python tools/gru_probe.py child 224 2>&1 | tail -2; done
timeout 100 python tools/layer_profile.py 224 5
