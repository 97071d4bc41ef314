# Round-2 final evidence (run under gpurun): plain run, ncu launch list of the same command, --set full of one batch-32 step,
# --set full of one impala step.  The .ncu-rep files stay in /tmp (gpurun_out/ is limited to 64 MiB).
set -x
B="python bench.py --steps 20 --warmup 3 --capacity 20000 --cpu-seconds 1 --no-dp --no-impala"
timeout 300 $B > gpurun_out/r02f_plain_b32.json 2> gpurun_out/r02f_plain_b32.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02f_bf16_launches.csv $B > gpurun_out/r02f_ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|reduce_seg|adam|heads_td|head_bwd|dense_fin|frames_to|gather_stack|sample_uniform" -s 64 -c 17 -o /tmp/r02f_b32_step -f $B > gpurun_out/r02f_ncu_full.log 2>&1
ncu -i /tmp/r02f_b32_step.ncu-rep --page raw --csv > gpurun_out/r02f_b32_step_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/r02f_impala -f python scripts/impala_ncu.py > gpurun_out/r02f_ncu_impala.log 2>&1
ncu -i /tmp/r02f_impala.ncu-rep --page raw --csv > gpurun_out/r02f_impala_raw.csv 2>/dev/null
tail -n 2 gpurun_out/r02f_ncu_full.log gpurun_out/r02f_ncu_impala.log | cut -c1-200; du -sh gpurun_out
