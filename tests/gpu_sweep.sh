set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2_t6_pytest.log
python tests/gpu_big_configs.py 2 3 4 > gpurun_out/r2_t6_base.log 2>&1
KSCHED_NO_WARPLOOP=1 python tests/gpu_big_configs.py 2 3 > gpurun_out/r2_t6_nowarp.log 2>&1
KSCHED_PACK_THREADS=64 python tests/gpu_big_configs.py 2 3 4 > gpurun_out/r2_t6_t64.log 2>&1
KSCHED_PACK_THREADS=32 python tests/gpu_big_configs.py 2 3 4 > gpurun_out/r2_t6_t32.log 2>&1
timeout 300 ncu --section WarpStateStats --section SourceCounters --section SchedulerStats --section InstructionStats --clock-control none --import-source on -k regex:pack_kernel -c 1 -o gpurun_out/r2_c4_pack4 python tests/gpu_big_configs.py 4 > gpurun_out/r2_t6_ncu.log 2>&1
