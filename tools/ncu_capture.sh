# ncu evidence for profiles/ (one GPU; every ncu pass runs only after the same command exited 0 without ncu)
#   1. launch list of bench.py itself (gpu__time_duration.sum)            -> gpurun_out/launches_bench_$R.csv
#   2. --set full of all 27 conv_gemm launches of one 888-segment call     -> gpurun_out/prof_gemm_$R.ncu-rep
#   3. --set full of the decoder GRU of the same call                      -> gpurun_out/prof_gru_$R.ncu-rep
set -x
R=${1:-r1h}
MB=${2:-888}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_$R.json 2> gpurun_out/bench_plain_$R.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$R.log 2>&1
python tools/profile_step.py $MB 2 > gpurun_out/plain_$R.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip 27 --launch-count 27 -f -o gpurun_out/prof_gemm_$R python tools/profile_step.py $MB 2 > gpurun_out/ncu_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_cluster_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_gru_$R python tools/profile_step.py $MB 2 >> gpurun_out/ncu_$R.log 2>&1
# summarise on the box (the 27-launch report is larger than gpurun_out may carry back) and keep only the CSVs
python tools/ncu_summary.py gpurun_out/prof_gemm_$R.ncu-rep gpurun_out/ncu_full_conv_gemm_$R.csv gemm > gpurun_out/gemm_traffic_$R.json
python tools/ncu_summary.py gpurun_out/prof_gru_$R.ncu-rep gpurun_out/ncu_full_gru_$R.csv > gpurun_out/gru_traffic_$R.json
rm -f gpurun_out/prof_gemm_$R.ncu-rep
tail -3 gpurun_out/ncu_$R.log
