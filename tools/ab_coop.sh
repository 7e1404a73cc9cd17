set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "msm or reference_shapes or pipelined or latency_path or degenerate or more_than_one or outer_msm or small_shapes" > gpurun_out/r02u_pytest_subset.log 2>&1; tail -3 gpurun_out/r02u_pytest_subset.log
B="--steps 5 --warmup 3 --no-single"
timeout 600 python bench.py $B --parity-sample 4 > gpurun_out/r02u_bench_prove.json 2> gpurun_out/r02u_bench_prove.err; tail -c 300 gpurun_out/r02u_bench_prove.err
MP_FWD_COOP=0 timeout 600 python bench.py $B --no-cpu-baseline > gpurun_out/r02u_bench_nocoop.json 2>/dev/null
python - <<'PY'
import json
for n in ('prove','nocoop'):
    try:
        d=json.load(open(f'gpurun_out/r02u_bench_{n}.json'))
        print(n, round(d['value'],1), round(d['e2e']['value'],1), str(d['parity'])[:40], {k.split('(')[0]:round(v,2) for k,v in d['phase_ms_per_step_serialised'].items()}, d['roofline']['ms_per_launch'])
    except Exception as e: print(n, 'ERR', e)
PY
