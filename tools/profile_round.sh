# Round profile: bench line, launch list, ncu full captures (run on the GPU box: gpurun -- 'bash tools/profile_round.sh')
set -x
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
B="python bench.py --steps 1 --warmup 1 --batch 32 --inflight 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_batch128.csv python bench.py --steps 1 --warmup 1 --batch 128 --inflight 1 --no-cpu-baseline > gpurun_out/ncu0.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 28 -c 1 -o gpurun_out/r01_ba_bwd_g1 -f $B > gpurun_out/ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 0 -c 1 -o gpurun_out/r01_ba_bwd_g2 -f $B > gpurun_out/ncu2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_fwd -s 28 -c 1 -o gpurun_out/r01_ba_fwd_g1 -f $B > gpurun_out/ncu3.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_ntt_cols|k_ntt_rows" -s 0 -c 2 -o gpurun_out/r01_ntt -f $B > gpurun_out/ncu4.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_mid -s 28 -c 1 -o gpurun_out/r01_ba_mid_g1 -f $B > gpurun_out/ncu5.log 2>&1
python tools/gpu_probe_extra.py --latency --poseidon --sweep --sweep-cpu > gpurun_out/probe_extra.log 2>&1
for f in r01_ba_bwd_g1 r01_ba_bwd_g2 r01_ba_fwd_g1 r01_ntt r01_ba_mid_g1; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; ncu -i gpurun_out/$f.ncu-rep --page details > gpurun_out/$f.details.txt 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep   # keep the merged output small: the csv / details exports are what profiles/ keeps
ls -la gpurun_out/
