set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02p_pytest_gpu.log
timeout 300 python bench.py --workload single --steps 50 > gpurun_out/r02p_bench_single.json 2>&1
timeout 300 python bench.py --workload single --steps 50 --no-cpu-baseline --shape to_private > gpurun_out/r02p_bench_single_tp.json 2>&1
