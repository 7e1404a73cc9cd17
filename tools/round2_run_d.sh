set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02d_pytest_gpu.log
timeout 600 python bench.py --parity-sample 4 --no-single > gpurun_out/r02d_bench_b128.json 2> gpurun_out/r02d_bench_b128.err; tail -c 300 gpurun_out/r02d_bench_b128.err
timeout 600 python bench.py --parity-sample 2 --no-single --batch 64 > gpurun_out/r02d_bench_b64.json 2>&1
timeout 600 python bench.py --parity-sample 2 --no-single --batch 32 > gpurun_out/r02d_bench_b32.json 2>&1
B="python bench.py --steps 1 --warmup 1 --batch 32 --inflight 1 --no-cpu-baseline --no-single"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02d_launches_batch128.csv python bench.py --steps 1 --warmup 1 --batch 128 --inflight 1 --no-cpu-baseline --no-single > gpurun_out/ncu0.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 28 -c 1 -o gpurun_out/r02d_ba_bwd_g1 -f $B > gpurun_out/ncu1.log 2>&1
for f in r02d_ba_bwd_g1; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; ncu -i gpurun_out/$f.ncu-rep --page details > gpurun_out/$f.details.txt 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep
