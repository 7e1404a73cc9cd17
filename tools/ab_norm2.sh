# kept norms (G2) + the 3-blocks-per-SM build of k_ba_bwd<Fq2>; outputs under gpurun_out/r02s_*
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "msm or reference_shapes or pipelined or latency_path or degenerate or more_than_one or outer_msm" > gpurun_out/r02s_pytest_subset.log 2>&1; tail -3 gpurun_out/r02s_pytest_subset.log
B="--steps 5 --warmup 3 --no-single --parity-sample 4"
timeout 600 python bench.py $B > gpurun_out/r02s_bench_prove.json 2> gpurun_out/r02s_bench_prove.err; tail -c 300 gpurun_out/r02s_bench_prove.err

timeout 300 python bench.py --workload g2_stress --no-cpu-baseline > gpurun_out/r02s_g2_stress.json 2>/dev/null

python - <<'PY'
import json
for n in ('',):
    try:
        d=json.load(open(f'gpurun_out/r02s_bench_prove{n}.json'))
        print('prove'+n, round(d['value'],1), round(d['e2e']['value'],1), d['parity'][:60], {k.split('(')[0]:round(v,2) for k,v in d['phase_ms_per_step_serialised'].items()}, d['device_bytes']['per_proof']>>20)
        d=json.load(open(f'gpurun_out/r02s_g2_stress{n}.json')); print('g2'+n, d.get('ms_per_step'), d.get('parity'))
    except Exception as e: print(n, 'ERR', e)
PY
