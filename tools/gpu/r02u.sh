set -x
python bench.py --workload walker_ppo --steps 5 --warmup 3 > gpurun_out/r02u_walker_ppo.json 2> gpurun_out/r02u_walker_ppo.err
python bench.py --workload cat_vf --steps 3 --warmup 3 > gpurun_out/r02u_cat_vf.json 2> gpurun_out/r02u_cat_vf.err
python bench.py --workload hopper --steps 20 --warmup 5 --verify full > gpurun_out/r02u_hopper.json 2> gpurun_out/r02u_hopper.err
# launch list of the headline bench
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --verify none > gpurun_out/r02u_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02u.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --verify none > gpurun_out/r02u_ncu_list.log 2>&1
# full captures of the small kernels
for k in cg_step_cluster_kernel reduce_partials_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_r02u_$k python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --verify none > gpurun_out/r02u_ncu_$k.log 2>&1
done
python bench.py --workload cat_vf --steps 1 --warmup 3 --verify none > gpurun_out/r02u_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gae_kernel -s 2 -c 1 -o gpurun_out/prof_r02u_gae_kernel python bench.py --workload cat_vf --steps 1 --warmup 3 --verify none > gpurun_out/r02u_ncu_gae.log 2>&1
tail -2 gpurun_out/r02u_*.err
