set -x
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
B="python bench.py --steps 1 --warmup 1 --batch 32 --inflight 1 --no-cpu-baseline"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 28 -c 1 -o gpurun_out/r01_ba_bwd_g1 -f $B > gpurun_out/ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 0 -c 1 -o gpurun_out/r01_ba_bwd_g2 -f $B > gpurun_out/ncu2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_fwd -s 28 -c 1 -o gpurun_out/r01_ba_fwd_g1 -f $B > gpurun_out/ncu3.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_ntt_cols|k_ntt_rows" -s 0 -c 2 -o gpurun_out/r01_ntt -f $B > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
