#!/bin/bash
# layer-1 kernels: two converter groups (default build) against one (tools/micro/libmrl_cg1.so)
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --verify quick --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" > gpurun_out/r02p_cg2.json
MRL_LIB=$PWD/tools/micro/libmrl_cg1.so python bench.py --steps 10 --warmup 3 --verify quick --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" > gpurun_out/r02p_cg1.json
python bench.py --steps 10 --warmup 3 --verify quick --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" > gpurun_out/r02p_cg2b.json
