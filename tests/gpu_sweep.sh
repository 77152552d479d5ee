set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2_t8_pytest.log
python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2_t8_bench.json 2> gpurun_out/r2_t8_bench.err
python bench.py --no-cpu-baseline --steps 5 --config 3 > gpurun_out/r2_t8_bench_c3.json 2>> gpurun_out/r2_t8_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^feasibility_kernel -c 1 -o gpurun_out/r2_k1_full python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2_t8_ncu.log 2>&1
