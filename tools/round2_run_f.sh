set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02f_pytest_gpu.log
timeout 300 python bench.py --workload single --steps 50 --no-cpu-baseline > gpurun_out/r02f_bench_single.json 2> gpurun_out/r02f_bench_single.err; tail -c 300 gpurun_out/r02f_bench_single.err
timeout 600 python bench.py --parity-sample 2 --no-single > gpurun_out/r02f_bench_tma.json 2> gpurun_out/r02f_bench_tma.err; tail -c 300 gpurun_out/r02f_bench_tma.err
MP_FWD_TMA=0 timeout 600 python bench.py --parity-sample 2 --no-single > gpurun_out/r02f_bench_notma.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02f_launches_tma.csv python bench.py --steps 1 --warmup 1 --batch 128 --inflight 1 --no-cpu-baseline --no-single > gpurun_out/ncu0.log 2>&1
