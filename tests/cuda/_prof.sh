set -x
for cfg in "1 128" "1 256" "2 256"; do
  set -- $cfg
  SPB_GEMM_NCTA=$1 SPB_GEMM_BN=$2 timeout 300 ncu --set full --clock-control none -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/gemm_n$1_bn$2 -f python tests/cuda/profile_gemm.py > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
