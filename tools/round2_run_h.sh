set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02h_pytest_gpu.log
timeout 600 python bench.py --parity-sample 4 > gpurun_out/r02h_bench_prove.json 2> gpurun_out/r02h_bench_prove.err; tail -c 300 gpurun_out/r02h_bench_prove.err
timeout 600 python bench.py --parity-sample 2 --no-single --batch 64 > gpurun_out/r02h_bench_b64.json 2>&1
timeout 600 python bench.py --parity-sample 2 --no-single --batch 32 > gpurun_out/r02h_bench_b32.json 2>&1
