#!/bin/bash
# per-role clock stamps of k_dense_chain / k_dense (variant built with -DGM_CHAIN_TRACE): where a layer's time goes inside
# a CTA.  Build the variant first (in racing_dreamer_b200/csrc, same flags as the Makefile):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden \
#        -DGM_CHAIN_TRACE -shared -o ../../variants/librd_env_chaintrace.so rd_env.cu -lcudart
OUT=gpurun_out/${1:-chaintrace}; mkdir -p $OUT
RD_ENV_LIB=variants/librd_env_chaintrace.so timeout 300 python - > $OUT/trace.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from racing_dreamer_b200 import BatchedRaceEnv, DreamerPolicy
wl = bench.workload_of(2)
env = BatchedRaceEnv(bench.env_config(wl, 4096, 0), device="cuda:0")
pol = DreamerPolicy(env, "austria_dreamer", noise="philox")
env.reset()
pol.rollout(45)
torch.cuda.synchronize()
PY
echo "rc=$?"; grep "CHAIN\|DENSE" $OUT/trace.txt | sort | head -120
