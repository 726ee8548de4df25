#!/bin/bash
python -m pytest tests -m gpu -q -k "multi or two_ranks" 2>&1 | tail -3
run() {  # name, port, timesteps
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --timesteps $3 --steps 10 --warmup 3 --verify quick --no-cpu-baseline 2> gpurun_out/r02m_$1.err | grep "^{" > gpurun_out/r02m_$1.json
}
run x250k 29541 250000
run x1m 29542 1000000
python bench.py --steps 10 --warmup 3 --verify quick --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r02m_1gpu.json
