# Round profile (run on the GPU box: gpurun -- 'bash tools/profile_round.sh'): every BASELINE config through bench.py, launch
# lists, ncu full captures of the dominant kernels, compute-sanitizer logs.  Outputs -> gpurun_out/r02_*; the summaries that
# are judged are copied to profiles/ afterwards (tools/ncu_summary.py).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_prove.json 2> gpurun_out/r02_bench_prove.err
timeout 300 python bench.py --workload single --steps 50 > gpurun_out/r02_bench_single.json 2> gpurun_out/r02_bench_single.err
timeout 900 python bench.py --workload msm_sweep --steps 5 > gpurun_out/r02_bench_msm_sweep.json 2> gpurun_out/r02_bench_msm_sweep.err
timeout 600 python bench.py --workload g2_stress --steps 5 > gpurun_out/r02_bench_g2_stress.json 2> gpurun_out/r02_bench_g2_stress.err
timeout 600 python bench.py --shape to_public --parity-sample 8 > gpurun_out/r02_bench_to_public.json 2> gpurun_out/r02_bench_to_public.err
timeout 600 python bench.py --shape to_private --parity-sample 8 > gpurun_out/r02_bench_to_private.json 2> gpurun_out/r02_bench_to_private.err
timeout 600 python bench.py --dist R --parity-sample 8 > gpurun_out/r02_bench_pt_distR.json 2> gpurun_out/r02_bench_pt_distR.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null
B="python bench.py --steps 1 --warmup 1 --batch 32 --inflight 1 --no-cpu-baseline --no-single"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_batch128.csv python bench.py --steps 1 --warmup 1 --batch 128 --inflight 1 --no-cpu-baseline --no-single > gpurun_out/ncu0.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_single.csv python bench.py --workload single --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_single.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 28 -c 1 -o gpurun_out/r02_ba_bwd_g1 -f $B > gpurun_out/ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 0 -c 1 -o gpurun_out/r02_ba_bwd_g2 -f $B > gpurun_out/ncu2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_fwd -s 28 -c 1 -o gpurun_out/r02_ba_fwd_g1 -f $B > gpurun_out/ncu3.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_ntt_cols|k_ntt_rows" -s 0 -c 2 -o gpurun_out/r02_ntt -f $B > gpurun_out/ncu4.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_mid -s 28 -c 1 -o gpurun_out/r02_ba_mid_g1 -f $B > gpurun_out/ncu5.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_prove_finish -s 1 -c 1 -o gpurun_out/r02_finish -f $B > gpurun_out/ncu6.log 2>&1
for f in r02_ba_bwd_g1 r02_ba_bwd_g2 r02_ba_fwd_g1 r02_ntt r02_ba_mid_g1 r02_finish; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; ncu -i gpurun_out/$f.ncu-rep --page details > gpurun_out/$f.details.txt 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep   # keep the merged output small: the csv / details exports are what profiles/ keeps
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_smoke.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_smoke.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02_sanitizer_racecheck.log
ls -la gpurun_out/
