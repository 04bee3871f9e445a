run() { echo "== $*"; "$@" 2>&1 | grep "M=" ; }
for envs in "" "UB_GEMM_NCTA=1" "UB_GEMM_BN=128"; do
  echo "#### env: $envs"
  for spec in "10240 768 768 bias=True resid=True out_fp32=1 time_it=True" "10240 768 768 b_mn=1 time_it=True" "768 768 10240 a_mn=1 b_mn=1 out_fp32=1 accumulate=1 split_k=8 time_it=True" "768 768 10240 a_mn=1 b_mn=1 out_fp32=1 accumulate=1 split_k=16 time_it=True" "10240 2304 768 bias=True time_it=True" "10240 768 2304 b_mn=1 time_it=True" "10240 768 3072 bias=True resid=True out_fp32=1 time_it=True" "10240 3072 768 bias=True act=2 time_it=True" "10240 512 768 bias=True out_fp32=1 time_it=True"; do
    env $envs python tools/gemm_check.py one $spec 2>&1 | grep "M="
  done
done
