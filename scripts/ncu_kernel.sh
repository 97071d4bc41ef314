# One `ncu --set full` capture of the kernels matching $1 in the short batch-32 bench; CSV pages land in gpurun_out/$2_*.csv
# usage: scripts/ncu_kernel.sh REGEX NAME [skip] [count] [ENV=...]
set -x
RE="$1"; NAME="$2"; SKIP="${3:-4}"; CNT="${4:-2}"; shift 4 || true
B="python bench.py --steps 20 --warmup 3 --capacity 20000 --cpu-seconds 1"
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o /tmp/$NAME -f $B > gpurun_out/ncu_$NAME.log 2>&1
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > gpurun_out/${NAME}_raw.csv 2>/dev/null
ncu -i /tmp/$NAME.ncu-rep --page details --csv > gpurun_out/${NAME}_details.csv 2>/dev/null
ncu -i /tmp/$NAME.ncu-rep --page source --csv > gpurun_out/${NAME}_source.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_$NAME.log
