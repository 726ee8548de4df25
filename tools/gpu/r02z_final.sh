#!/bin/bash
# round-2 final 1-GPU record: tests, ncu captures of the three tensor-core kernels at 1M timesteps (DRAM traffic for
# roofline.traffic), the headline bench with full verification, the reference arm, the launch list
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02z_pytest.log
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --verify none"
$B > gpurun_out/r02z_plain.log 2>&1 || exit 1
for k in fvp_tc_kernel l1_forward_tc_kernel l1_grad_tc_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o gpurun_out/prof_r02z_$k $B > gpurun_out/r02z_ncu_$k.log 2>&1
done
python tools/ncu_summary.py traffic profiles/ncu_traffic.json mid_backward_fvp=gpurun_out/prof_r02z_fvp_tc_kernel.ncu-rep \
  l1_forward=gpurun_out/prof_r02z_l1_forward_tc_kernel.ncu-rep l1_grad=gpurun_out/prof_r02z_l1_grad_tc_kernel.ncu-rep > gpurun_out/r02z_traffic.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/r02z_ncu_traffic.json
MRL_WRITE_N1_STATS=1 python bench.py --steps 20 --warmup 5 --verify full > gpurun_out/r02z_bench_1gpu.json 2> gpurun_out/r02z_bench_1gpu.err
cp profiles/n1_stats_*.json gpurun_out/
python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 60 > gpurun_out/r02z_ref.json 2> gpurun_out/r02z_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02z.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --verify none > gpurun_out/r02z_ncu_list.log 2>&1
tail -3 gpurun_out/r02z_pytest.log
