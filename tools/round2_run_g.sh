set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02g_pytest_gpu.log
timeout 300 python bench.py --workload single --steps 50 --no-cpu-baseline > gpurun_out/r02g_bench_single.json 2> gpurun_out/r02g_bench_single.err; tail -c 300 gpurun_out/r02g_bench_single.err
timeout 600 python bench.py --parity-sample 2 --no-single > gpurun_out/r02g_bench_prove.json 2> gpurun_out/r02g_bench_prove.err; tail -c 300 gpurun_out/r02g_bench_prove.err
timeout 300 python bench.py --workload msm_sweep --max-log 20 --steps 5 --no-cpu-baseline > gpurun_out/r02g_bench_msm_sweep.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02g_launches_single.csv python bench.py --workload single --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_single.log 2>&1
