# Round-2 evidence (run under gpurun): GPU tests of what changed last, then ncu captures exported to CSV pages on the box
# (the .ncu-rep files stay in /tmp: gpurun_out/ is limited to 64 MiB).
set -x
timeout 900 python -m pytest tests/test_acting_gpu.py tests/test_samplers_gpu.py tests/test_sum_tree_gpu.py tests/test_replay_batch_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_pytest7.log 2>&1; tail -n 3 gpurun_out/r2_pytest7.log
B="python bench.py --steps 20 --warmup 3 --capacity 20000 --cpu-seconds 1 --no-dp"
timeout 300 $B > gpurun_out/r02_plain_b32.json 2> gpurun_out/r02_plain_b32.err || exit 1
# (1) launch list of the whole short bench
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_bf16_launches.csv $B > gpurun_out/r02_ncu_ll.log 2>&1
# (2) every kernel of one batch-32 step, --set full
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|reduce_seg|adam|heads_td|head_bwd|dense_fin|frames_to|gather_stack|sample_uniform" -s 64 -c 17 -o /tmp/r02_b32_step -f $B > gpurun_out/r02_ncu_full.log 2>&1
ncu -i /tmp/r02_b32_step.ncu-rep --page raw --csv > gpurun_out/r02_b32_step_raw.csv 2>/dev/null
# (3) replay kernels at their throughput shapes (profiling starts after the fills)
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"gather_stack4|sample_prioritized|sumtree_set|sumtree_keys|uniform_|sumtree_query" -c 24 -o /tmp/r02_replay -f python scripts/replay_kernels.py > gpurun_out/r02_ncu_replay.log 2>&1
ncu -i /tmp/r02_replay.ncu-rep --page raw --csv > gpurun_out/r02_replay_raw.csv 2>/dev/null
# (4) the tensor-core kernels at batch 4096, widths x1 and x4
for W in 1 4; do
  D="python bench.py --mode dp --batch 4096 --width $W --steps 2 --warmup 3 --capacity 20000"
  timeout 300 $D > gpurun_out/r02_plain_dp_w$W.json 2> gpurun_out/r02_plain_dp_w$W.err
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|frames_to" -s 36 -c 12 -o /tmp/r02_dp4096_w$W -f $D > gpurun_out/r02_ncu_dp_w$W.log 2>&1
  ncu -i /tmp/r02_dp4096_w$W.ncu-rep --page raw --csv > gpurun_out/r02_dp4096_w${W}_raw.csv 2>/dev/null
done
tail -n 2 gpurun_out/r02_ncu_full.log gpurun_out/r02_ncu_replay.log gpurun_out/r02_ncu_dp_w1.log gpurun_out/r02_ncu_dp_w4.log | cut -c1-200; du -sh gpurun_out
