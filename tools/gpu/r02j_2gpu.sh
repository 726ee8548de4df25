#!/bin/bash
# 2 GPUs x 125k timesteps (the per-GPU shape of the 8-GPU run): push (round-1 transport) against pull
run() {  # name, extra env
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --timesteps 250000 --steps 10 --warmup 3 --verify none --no-cpu-baseline 2> gpurun_out/r02j_$1.err | grep "^{" > gpurun_out/r02j_$1.json
}
run pull "A=1" 29541
run push "MRL_LIB=$PWD/tools/micro/libmrl_push.so" 29542
run pull2 "A=1" 29543
