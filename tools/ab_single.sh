# A/B of the single-proof latency path on one B200 (switch names as of the final tree; the r02q_* artifacts were taken when the MSM form of the ladders and two trimmed levels were the defaults); outputs under gpurun_out/r02q_*
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02q_pytest_gpu.log
S="--workload single --no-cpu-baseline"
MP_LADDERS_AS_MSM=1 MP_BA_TRIM_LEVELS=2 timeout 300 python bench.py $S > gpurun_out/r02q_single_new.json 2> gpurun_out/r02q_single_new.err; tail -c 300 gpurun_out/r02q_single_new.err
MP_SMALL_PATH_OLD=1 MP_BA_TRIM_LEVELS=0 timeout 300 python bench.py $S > gpurun_out/r02q_single_old.json 2>/dev/null
MP_BA_TRIM_LEVELS=2 timeout 300 python bench.py $S > gpurun_out/r02q_single_ladders.json 2>/dev/null
MP_LADDERS_AS_MSM=1 MP_BA_TRIM_LEVELS=0 timeout 300 python bench.py $S > gpurun_out/r02q_single_trim0.json 2>/dev/null
MP_LADDERS_AS_MSM=1 MP_BA_TRIM_LEVELS=1 timeout 300 python bench.py $S > gpurun_out/r02q_single_trim1.json 2>/dev/null
MP_LADDERS_AS_MSM=1 MP_BA_TRIM_LEVELS=3 timeout 300 python bench.py $S > gpurun_out/r02q_single_trim3.json 2>/dev/null
timeout 300 python bench.py --workload msm_sweep --no-cpu-baseline --max-log 20 > gpurun_out/r02q_msm_sweep.json 2>/dev/null
for f in gpurun_out/r02q_single_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), d.get("parity","")[:60], {k.split('(')[0]:round(v,2) for k,v in d.get("phases_ms_last",d.get("single_proof",{}).get("phases_ms_last",{})).items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
