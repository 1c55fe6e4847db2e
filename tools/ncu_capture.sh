# ncu evidence for profiles/: launch list + --set full captures of the dominant kernels (one GPU, after a plain run exits 0)
set -x
R=${1:-r1g}
MB=${2:-222}
python tools/profile_step.py $MB 2 > gpurun_out/plain_$R.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python tools/profile_step.py $MB 2 > gpurun_out/ncu_$R.log 2>&1
# every conv_gemm launch of the second repetition (27 per encode->decode)
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip 27 --launch-count 27 -f -o gpurun_out/prof_gemm_$R python tools/profile_step.py $MB 2 >> gpurun_out/ncu_$R.log 2>&1
# the decoder GRU of the second repetition (launch order: enc, dec, enc, dec)
ncu --set full --clock-control none --import-source on -k regex:gru_cluster_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_gru_$R python tools/profile_step.py $MB 2 >> gpurun_out/ncu_$R.log 2>&1
tail -3 gpurun_out/ncu_$R.log
