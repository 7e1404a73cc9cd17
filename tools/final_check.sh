# last check of the round on one B200: GPU tests, smoke, the default bench line, the reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02z_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; tail -2 gpurun_out/r02z_smoke.log
timeout 600 python bench.py > gpurun_out/r02z_bench_prove_final.json 2> gpurun_out/r02z_bench_prove_final.err; tail -c 300 gpurun_out/r02z_bench_prove_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench_prove_final.json'))
print(round(d['value'],1), round(d['e2e']['value'],1), d['parity'][:60], d['clocks'], d['device_bytes'], d['single_proof']['e2e_ms_median'], d['single_proof'].get('device_bytes'))
PY
