# Round-end evidence: GPU tests, plain benches, ncu launch lists and one --set full capture of a batch-32 step and of the
# tensor-core kernels at batch 4096.  Run under gpurun; the .ncu-rep files are exported to CSV pages on the box and
# removed (gpurun_out/ is limited to 64 MiB), everything else lands in gpurun_out/.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
B="python bench.py --steps 20 --warmup 3 --capacity 20000 --cpu-seconds 1"
timeout 300 $B > gpurun_out/plain_b32.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_bf16_launches.csv $B > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|ln_relu|reduce_seg|adam|heads_td|head_bwd|dense_fin|frames_to|act_forward" -s 63 -c 21 -o /tmp/r01_b32_step -f $B > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/r01_b32_step.ncu-rep --page raw --csv > gpurun_out/r01_b32_step_raw.csv 2>/dev/null
ncu -i /tmp/r01_b32_step.ncu-rep --page details --csv > gpurun_out/r01_b32_step_details.csv 2>/dev/null
D="python bench.py --mode dp --batch 4096 --width 1 --steps 2 --warmup 3"
timeout 300 $D > gpurun_out/plain_dp.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_dp4096_launches.csv $D > gpurun_out/ncu_ll_dp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k tc_gemm_kernel -s 33 -c 11 -o /tmp/r01_dp4096_tc -f $D > gpurun_out/ncu_full_dp.log 2>&1
ncu -i /tmp/r01_dp4096_tc.ncu-rep --page raw --csv > gpurun_out/r01_dp4096_tc_raw.csv 2>/dev/null
ncu -i /tmp/r01_dp4096_tc.ncu-rep --page details --csv > gpurun_out/r01_dp4096_tc_details.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_full.log; tail -n 2 gpurun_out/ncu_full_dp.log; du -sh gpurun_out
