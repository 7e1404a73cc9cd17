set -x
mkdir -p gpurun_out
MP_G2_BWD_SMEM=1 timeout 900 python -m pytest tests -m gpu -x -q -k "msm or reference_shapes or pipelined or latency_path or degenerate or more_than_one or outer_msm or small_shapes" > gpurun_out/r02v_pytest_subset.log 2>&1; tail -3 gpurun_out/r02v_pytest_subset.log
B="--steps 5 --warmup 3 --no-single"
MP_G2_BWD_SMEM=1 timeout 600 python bench.py $B --parity-sample 4 > gpurun_out/r02v_bench_smem.json 2> gpurun_out/r02v_bench_smem.err; tail -c 300 gpurun_out/r02v_bench_smem.err
timeout 600 python bench.py $B --no-cpu-baseline > gpurun_out/r02v_bench_regs.json 2>/dev/null
MP_G2_BWD_SMEM=1 timeout 300 python bench.py --workload g2_stress --no-cpu-baseline > gpurun_out/r02v_g2_stress_smem.json 2>/dev/null
python - <<'PY'
import json
for n in ('smem','regs'):
    try:
        d=json.load(open(f'gpurun_out/r02v_bench_{n}.json'))
        print(n, round(d['value'],1), round(d['e2e']['value'],1), str(d['parity'])[:40], {k.split('(')[0]:round(v,2) for k,v in d['phase_ms_per_step_serialised'].items()}, d['roofline']['ms_per_launch'])
    except Exception as e: print(n, 'ERR', e)
d=json.load(open('gpurun_out/r02v_g2_stress_smem.json')); print('g2 smem', d.get('ms_per_step'))
PY
