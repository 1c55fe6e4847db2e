# ncu --set full + source of selected conv_gemm launches of one 960-segment call (launch index within the call's 27 GEMM launches):
#   sh tools/ncu_pick.sh TAG idx [idx ...]     -> gpurun_out/prof_TAG_<idx>.ncu-rep
TAG=$1; shift
timeout 120 python tools/profile_step.py 960 2 > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
for i in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip $((27 + i)) --launch-count 1 -f -o gpurun_out/prof_${TAG}_$i python tools/profile_step.py 960 2 > gpurun_out/${TAG}_ncu_$i.log 2>&1
done
ls -la gpurun_out/prof_${TAG}_*
