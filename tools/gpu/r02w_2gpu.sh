#!/bin/bash
timeout 300 python -m pytest tests -m gpu -q -k "two_ranks or multi" 2>&1 | tail -2
run() {  # name, port, timesteps, env
  env $4 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --timesteps $3 --steps 20 --warmup 5 --verify quick --no-cpu-baseline 2> gpurun_out/r02w_$1.err | grep "^{" > gpurun_out/r02w_$1.json
}
run x1m 29542 1000000 A=1
run x1m_split 29543 1000000 MRL_P2P_SPLIT=1
run x250k 29541 250000 A=1
