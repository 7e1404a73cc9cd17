# base-field shared inversion for G2 + final defaults; outputs under gpurun_out/r02r_*
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02r_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02r_bench_prove.json 2> gpurun_out/r02r_bench_prove.err; tail -c 300 gpurun_out/r02r_bench_prove.err
S="--workload single --no-cpu-baseline"
timeout 300 python bench.py $S > gpurun_out/r02r_single_neworder.json 2>/dev/null
MP_SMALL_PATH_OLD=1 timeout 300 python bench.py $S > gpurun_out/r02r_single_oldorder.json 2>/dev/null
timeout 300 python bench.py --workload g2_stress --no-cpu-baseline > gpurun_out/r02r_g2_stress.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02r_bench_prove.json'))
print('prove', round(d['value'],1), round(d['e2e']['value'],1), d['parity'][:60], {k.split('(')[0]:round(v,2) for k,v in d['phase_ms_per_step_serialised'].items()}, d['device_bytes']['per_proof']>>20, d['single_proof']['e2e_ms_median'])
for n in ('neworder','oldorder'):
    d=json.load(open(f'gpurun_out/r02r_single_{n}.json')); print(n, d['latency_ms'])
d=json.load(open('gpurun_out/r02r_g2_stress.json')); print('g2', d.get('ms_per_step'), d.get('value'))
PY
