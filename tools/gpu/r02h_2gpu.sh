#!/bin/bash
# 2-GPU bench with reduce-kernel grid variants
for cps in 6 2 8; do
  MRL_RED_CPS=$cps python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520+cps)) bench.py --gpus 2 --steps 10 --warmup 3 --verify quick 2> gpurun_out/r02h_2gpu_cps$cps.err | grep "^{" > gpurun_out/r02h_2gpu_cps$cps.json
done
python bench.py --steps 10 --warmup 3 --verify quick 2>/dev/null | grep "^{" > gpurun_out/r02h_1gpu.json
