set -x
R=r1j
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_$R.json 2> gpurun_out/bench_plain_$R.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$R.log 2>&1
python tools/profile_step.py 960 2 > gpurun_out/plain_$R.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gru_wide_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_gru_$R python tools/profile_step.py 960 2 > gpurun_out/ncu_$R.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_gru_$R.ncu-rep gpurun_out/ncu_full_gru_$R.csv > gpurun_out/gru_traffic_$R.json
rm -f gpurun_out/prof_gru_$R.ncu-rep
tail -2 gpurun_out/ncu_$R.log
