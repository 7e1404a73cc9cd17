# multi-GPU pass: bash tools/round2_run_multi.sh N   (N = number of GPUs of the box)
N=${1:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_prove_${N}gpu.json 2> gpurun_out/r02_bench_prove_${N}gpu.err; tail -c 500 gpurun_out/r02_bench_prove_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --workload g2_stress --steps 5 > gpurun_out/r02_bench_g2_stress_${N}gpu.json 2> gpurun_out/r02_bench_g2_stress_${N}gpu.err; tail -c 300 gpurun_out/r02_bench_g2_stress_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --shape to_public --parity-sample 8 --no-single > gpurun_out/r02_bench_to_public_${N}gpu.json 2> gpurun_out/r02_bench_to_public_${N}gpu.err
timeout 300 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_${N}gpu.json 2>/dev/null
grep -h -o '"value": [0-9.]*' gpurun_out/r02_bench_*_${N}gpu.json | head
