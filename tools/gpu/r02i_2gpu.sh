#!/bin/bash
# 2-GPU: multi-GPU tests + bench with the pull transport
python -m pytest tests -m gpu -q -k "multi or comm or parallel or rank" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --verify quick 2> gpurun_out/r02i_2gpu.err | grep "^{" > gpurun_out/r02i_2gpu.json
tail -3 gpurun_out/r02i_2gpu.err
python bench.py --steps 10 --warmup 3 --verify quick 2>/dev/null | grep "^{" > gpurun_out/r02i_1gpu.json
