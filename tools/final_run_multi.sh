# multi-GPU pass on the final tree: bash tools/final_run_multi.sh N [full]   (N = number of GPUs of the box)
N=${1:-8}
set -x
mkdir -p gpurun_out
P=gpurun_out/r02z
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-single > ${P}_bench_prove_${N}gpu.json 2> ${P}_bench_prove_${N}gpu.err; tail -c 300 ${P}_bench_prove_${N}gpu.err
if [ "$2" = "full" ]; then
timeout 600 $TR bench.py --gpus $N --workload g2_stress --steps 5 > ${P}_bench_g2_stress_${N}gpu.json 2> ${P}_bench_g2_stress_${N}gpu.err; tail -c 300 ${P}_bench_g2_stress_${N}gpu.err
timeout 300 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > ${P}_bench_reference_${N}gpu.json 2>/dev/null
fi
grep -h -o '"value": [0-9.]*' ${P}_bench_*_${N}gpu.json | head
grep -h -o '"parity": "[^"]*"' ${P}_bench_*_${N}gpu.json | head
