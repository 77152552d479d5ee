set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2_t7_pytest.log
python bench.py > gpurun_out/r2_t7_bench.json 2> gpurun_out/r2_t7_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_t7_ref.json 2>&1
