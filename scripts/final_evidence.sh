# Round-end evidence on ONE box: GPU tests, the default bench line, the large-batch lines, acting breakdown, ncu launch
# list + one --set full capture of the batch-32 step (scripts/run_ncu.sh does the heavy part).
set -x
bash scripts/run_ncu.sh
python scripts/acting_profile.py > gpurun_out/acting_profile.txt 2>&1
for w in 1 2 4; do
  timeout 300 python bench.py --mode dp --batch 4096 --width $w --steps 10 --warmup 3 > gpurun_out/dp_w$w.json 2> gpurun_out/dp_w$w.err || tail -n 3 gpurun_out/dp_w$w.err
done
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || tail -n 5 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 600 gpurun_out/bench_default.json; du -sh gpurun_out
