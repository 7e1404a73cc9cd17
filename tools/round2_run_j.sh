set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02j_pytest_gpu.log
timeout 600 python bench.py --parity-sample 4 --no-single > gpurun_out/r02j_bench_prove.json 2> gpurun_out/r02j_bench_prove.err; tail -c 300 gpurun_out/r02j_bench_prove.err
MP_MSM_XYZZ=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prove_small or msm_vs_oracle or golden or batch_api" > gpurun_out/r02j_pytest_xyzz.log 2>&1; tail -3 gpurun_out/r02j_pytest_xyzz.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02j_launches_batch128.csv python bench.py --steps 1 --warmup 1 --batch 128 --inflight 1 --no-cpu-baseline --no-single > gpurun_out/ncu0.log 2>&1
