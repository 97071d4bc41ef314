set -x
D="python bench.py --mode dp --batch 4096 --width 1 --steps 2 --warmup 3"
timeout 300 $D > gpurun_out/plain_dp.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k tc_gemm_kernel -s 33 -c 11 -o gpurun_out/r01_dp4096_tc -f $D > gpurun_out/ncu_full_dp.log 2>&1
tail -3 gpurun_out/ncu_full_dp.log | cut -c1-300
