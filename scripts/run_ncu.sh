set -x
B="python bench.py --steps 20 --warmup 3 --capacity 20000 --cpu-seconds 1"
timeout 300 $B > gpurun_out/plain_b32.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_bf16_launches.csv $B > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|ln_relu|reduce_seg|adam|heads_td|gemm_strided|dense_fin|head_fwd" -s 66 -c 22 -o gpurun_out/r01_b32_step -f $B > gpurun_out/ncu_full.log 2>&1
D="python bench.py --mode dp --batch 4096 --width 1 --steps 2 --warmup 3"
timeout 300 $D > gpurun_out/plain_dp.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_dp4096_launches.csv $D > gpurun_out/ncu_ll_dp.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
