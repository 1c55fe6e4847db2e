# ncu evidence for profiles/ (round 2, final kernels; one GPU; every ncu pass runs only after the same command exited 0 without ncu)
#   1. launch list of bench.py itself (gpu__time_duration.sum)                      -> gpurun_out/r02_launches_bench.csv
#   2. --set full of the 27 conv_gemm launches of one 960-segment call, summarised  -> gpurun_out/r02_ncu_full_conv_gemm_mb960.csv
#   3. --set full of the decoder GRU, summarised                                     -> gpurun_out/r02_ncu_full_gru_dec_mb960.csv
set -x
timeout 300 python bench.py --steps 2 --warmup 3 --calls-per-step 2 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --calls-per-step 2 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
timeout 120 python tools/profile_step.py 960 2 > gpurun_out/r02_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip 27 --launch-count 27 -f -o gpurun_out/prof_gemm_r02 python tools/profile_step.py 960 2 > gpurun_out/r02_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gru_wide_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_gru_r02 python tools/profile_step.py 960 2 >> gpurun_out/r02_ncu_gemm.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_gemm_r02.ncu-rep gpurun_out/r02_ncu_full_conv_gemm_mb960.csv gemm | python -c "import json,sys; d=json.load(sys.stdin); d['micro_batch']=960; print(json.dumps(d))" > gpurun_out/r02_gemm_traffic.json
python tools/ncu_summary.py gpurun_out/prof_gru_r02.ncu-rep gpurun_out/r02_ncu_full_gru_dec_mb960.csv > gpurun_out/r02_gru_traffic.json
rm -f gpurun_out/prof_gemm_r02.ncu-rep gpurun_out/prof_gru_r02.ncu-rep
tail -3 gpurun_out/r02_ncu_gemm.log
