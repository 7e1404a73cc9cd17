# k_ba_mid grids rounded down to whole waves; outputs under gpurun_out/r02w_*
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "pipelined or more_than_one or outer_msm or msm_vs_oracle or closed_form_large" > gpurun_out/r02w_pytest_subset.log 2>&1; tail -2 gpurun_out/r02w_pytest_subset.log
B="--steps 5 --warmup 3 --no-single"
timeout 300 python bench.py $B --parity-sample 2 > gpurun_out/r02w_bench_midwave.json 2> gpurun_out/r02w_bench_midwave.err; tail -c 200 gpurun_out/r02w_bench_midwave.err
MP_BA_MID_WAVE=0 timeout 300 python bench.py $B --no-cpu-baseline > gpurun_out/r02w_bench_plain.json 2>/dev/null
python - <<'PY'
import json
for n in ('midwave','plain'):
    try:
        d=json.load(open(f'gpurun_out/r02w_bench_{n}.json'))
        print(n, round(d['value'],1), round(d['e2e']['value'],1), str(d['parity'])[:40], {k.split('(')[0]:round(v,2) for k,v in d['phase_ms_per_step_serialised'].items()})
    except Exception as e: print(n, 'ERR', e)
PY
