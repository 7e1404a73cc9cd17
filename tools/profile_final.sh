# Final-tree profile of round 2 (gpurun -- 'bash tools/profile_final.sh'): every BASELINE config through bench.py, launch lists,
# ncu full captures of the kernels that changed since tools/profile_round.sh ran (G2 rounds share their inversions in Fq),
# compute-sanitizer memcheck.  Outputs -> gpurun_out/r02z_*; summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
P=gpurun_out/r02z
timeout 1200 python -m pytest tests -m gpu -x -q > ${P}_pytest_gpu.log 2>&1; tail -3 ${P}_pytest_gpu.log
timeout 600 python bench.py > ${P}_bench_prove.json 2> ${P}_bench_prove.err
timeout 300 python bench.py --workload single --steps 50 > ${P}_bench_single.json 2> ${P}_bench_single.err
timeout 300 python bench.py --workload single --steps 50 --shape to_private --no-cpu-baseline > ${P}_bench_single_to_private.json 2>/dev/null
timeout 900 python bench.py --workload msm_sweep --steps 5 > ${P}_bench_msm_sweep.json 2> ${P}_bench_msm_sweep.err
timeout 600 python bench.py --workload g2_stress --steps 5 > ${P}_bench_g2_stress.json 2> ${P}_bench_g2_stress.err
timeout 600 python bench.py --shape to_public --parity-sample 8 > ${P}_bench_to_public.json 2> ${P}_bench_to_public.err
timeout 600 python bench.py --shape to_private --parity-sample 8 > ${P}_bench_to_private.json 2> ${P}_bench_to_private.err
timeout 600 python bench.py --dist R --parity-sample 8 > ${P}_bench_pt_distR.json 2> ${P}_bench_pt_distR.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2>/dev/null
B="python bench.py --steps 1 --warmup 1 --batch 32 --inflight 1 --no-cpu-baseline --no-single"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${P}_launches_batch128.csv python bench.py --steps 1 --warmup 1 --batch 128 --inflight 1 --no-cpu-baseline --no-single > gpurun_out/ncu0.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${P}_launches_single.csv python bench.py --workload single --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_single.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 28 -c 1 -o ${P}_ba_bwd_g1 -f $B > gpurun_out/ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 0 -c 1 -o ${P}_ba_bwd_g2 -f $B > gpurun_out/ncu2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_fwd -s 0 -c 1 -o ${P}_ba_fwd_g2 -f $B > gpurun_out/ncu3.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_mid -s 0 -c 1 -o ${P}_ba_mid_g2 -f $B > gpurun_out/ncu5.log 2>&1
for f in ba_bwd_g1 ba_bwd_g2 ba_fwd_g2 ba_mid_g2; do ncu -i ${P}_$f.ncu-rep --page raw --csv > ${P}_$f.raw.csv 2>/dev/null; ncu -i ${P}_$f.ncu-rep --page details > ${P}_$f.details.txt 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_smoke.py > ${P}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> ${P}_sanitizer_memcheck.log
ls -la gpurun_out/ | grep r02z
