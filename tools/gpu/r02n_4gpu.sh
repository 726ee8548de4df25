#!/bin/bash
run() {  # name, port, env
  env $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 4 --steps 20 --warmup 5 --verify quick --no-cpu-baseline 2> gpurun_out/r02n_$1.err | grep "^{" > gpurun_out/r02n_$1.json
}
run 4gpu 29541 A=1
run 4gpu_bal 29542 MRL_RED_BALANCE=1
