set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02e_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02e_launches_single.csv python bench.py --workload single --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_single.log 2>&1
