# A/B of the slab pipeline of the tree levels (accumulate_ba) on one B200; outputs under gpurun_out/r02p_*
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pipelined or outer_msm or more_than_one or chunked or small_shapes" > gpurun_out/r02p_pytest_subset.log 2>&1; tail -3 gpurun_out/r02p_pytest_subset.log
B="--steps 5 --warmup 3 --no-single"
timeout 600 python bench.py $B > gpurun_out/r02p_bench_pipe_div4.json 2> gpurun_out/r02p_bench_pipe_div4.err; tail -c 200 gpurun_out/r02p_bench_pipe_div4.err
MP_BA_PIPE=0 timeout 600 python bench.py $B --no-cpu-baseline > gpurun_out/r02p_bench_nopipe.json 2>/dev/null
MP_BA_SLAB_DIV=2 timeout 600 python bench.py $B --no-cpu-baseline > gpurun_out/r02p_bench_pipe_div2.json 2>/dev/null
MP_BA_PIPE_MIN_LOG2=40 timeout 600 python bench.py $B --no-cpu-baseline > gpurun_out/r02p_bench_pipe_div4_noforce.json 2>/dev/null
MP_BA_PIPE_MIN_LOG2=25 timeout 600 python bench.py $B --no-cpu-baseline > gpurun_out/r02p_bench_pipe_div4_min25.json 2>/dev/null
for f in gpurun_out/r02p_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), d.get("parity","")[:40], d["device_bytes"]["per_proof"]>>20, d["roofline"]["frac"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
