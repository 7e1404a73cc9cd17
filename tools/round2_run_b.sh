# Round-2 GPU pass B: new finishing kernel + small-batch reduction
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02b_pytest_gpu.log
timeout 300 python bench.py --workload single --steps 50 > gpurun_out/r02b_bench_single.json 2> gpurun_out/r02b_bench_single.err; tail -c 400 gpurun_out/r02b_bench_single.err
MP_NO_GLV=1 timeout 300 python bench.py --workload single --steps 50 --no-cpu-baseline > gpurun_out/r02b_bench_single_noglv.json 2>&1
timeout 600 python bench.py --parity-sample 4 > gpurun_out/r02b_bench_prove.json 2> gpurun_out/r02b_bench_prove.err; tail -c 400 gpurun_out/r02b_bench_prove.err
timeout 300 python bench.py --workload msm_sweep --max-log 20 --steps 5 > gpurun_out/r02b_bench_msm_sweep.json 2> gpurun_out/r02b_bench_msm_sweep.err
