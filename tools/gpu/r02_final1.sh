set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02f_pytest.log
MRL_WRITE_N1_STATS=1 python bench.py --steps 20 --warmup 5 --verify full > gpurun_out/r02f_bench_1gpu.json 2> gpurun_out/r02f_bench_1gpu.err
cp profiles/n1_stats_*.json gpurun_out/
python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 60 > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err
python tools/micro/zf_run.py > gpurun_out/r02f_zf_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:zf_ -c 4 -o gpurun_out/prof_r02f_zf python tools/micro/zf_run.py > gpurun_out/r02f_zf_ncu.log 2>&1
tail -3 gpurun_out/r02f_pytest.log
