#!/bin/bash
# layer-1 kernels: converter groups / k-groups per stage
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "synthetic_shapes" 2>&1 | tail -3
b() { env $2 python bench.py --steps 6 --warmup 3 --verify quick --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" > gpurun_out/r02q_$1.json; }
b cg2 A=1
b cg1 MRL_L1_CGROUPS=1
b cg2_kps2 MRL_L1_KPS=2
b cg2_kps4 MRL_L1_KPS=4
