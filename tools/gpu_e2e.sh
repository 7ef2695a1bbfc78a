#!/bin/bash
# GPU tests + bench + PCIe probe (+ 2-GPU torchrun bench when 2 GPUs are visible)
OUT=gpurun_out/${1:-e2e}; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
python - <<'PY'
import torch, time
n=64<<20
d=torch.empty(n,dtype=torch.uint8,device='cuda'); h=torch.empty(n,dtype=torch.uint8,pin_memory=True)
for name,(dst,src) in {'d2h':(h,d),'h2d':(d,h)}.items():
    for _ in range(3): dst.copy_(src,non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): dst.copy_(src,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(f"pcie {name} {10*n/dt/1e9:.1f} GB/s (64 MiB pinned)")
PY
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 50 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "bench2 rc=$?"; cat $OUT/bench_2gpu.json; tail -5 $OUT/bench_2gpu.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $OUT/bench_2gpu_ref.json 2> $OUT/bench_2gpu_ref.err; echo "bench2ref rc=$?"; cat $OUT/bench_2gpu_ref.json
fi
