#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:"net_gemm_kernel|splitk_reduce_kernel" -c 6 -o gpurun_out/r2j_gemm python -c "
import sys; sys.path.insert(0,'.')
import torch, bench
torch.cuda.profiler.start()
print(bench.offpolicy_extras(torch.device('cuda:0')).keys())
torch.cuda.profiler.stop()" > gpurun_out/r2j_ncu_gemm.log 2>&1
tail -3 gpurun_out/r2j_ncu_gemm.log
