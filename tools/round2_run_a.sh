# Round-2 GPU pass A: parity tests on the reworked library, then one line per BASELINE config (outputs -> gpurun_out/)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02a_bench_prove.json 2> gpurun_out/r02a_bench_prove.err; tail -c 600 gpurun_out/r02a_bench_prove.err
timeout 300 python bench.py --workload single --steps 50 > gpurun_out/r02a_bench_single.json 2> gpurun_out/r02a_bench_single.err
MP_MSM_XYZZ=1 timeout 300 python bench.py --workload single --steps 50 > gpurun_out/r02a_bench_single_xyzz.json 2> gpurun_out/r02a_bench_single_xyzz.err
timeout 900 python bench.py --workload msm_sweep --steps 5 > gpurun_out/r02a_bench_msm_sweep.json 2> gpurun_out/r02a_bench_msm_sweep.err
timeout 600 python bench.py --workload g2_stress --steps 5 > gpurun_out/r02a_bench_g2_stress.json 2> gpurun_out/r02a_bench_g2_stress.err
timeout 600 python bench.py --shape to_public --parity-sample 8 > gpurun_out/r02a_bench_to_public.json 2> gpurun_out/r02a_bench_to_public.err
timeout 600 python bench.py --dist R --parity-sample 8 > gpurun_out/r02a_bench_pt_distR.json 2> gpurun_out/r02a_bench_pt_distR.err
ls -la gpurun_out | tail -20
