#!/bin/bash
run() {  # name, env, port, timesteps
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --timesteps $4 --steps 10 --warmup 3 --verify quick --no-cpu-baseline 2> gpurun_out/r02k_$1.err | grep "^{" > gpurun_out/r02k_$1.json
}
run pull250k "A=1" 29541 250000
run pull1m "A=1" 29542 1000000
python bench.py --steps 10 --warmup 3 --verify quick --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r02k_1gpu.json
