# A/B of environment switches on the batch-32 bench (short, small replay): scripts/ab_bench.sh NAME=ENVSPEC ...
# e.g. scripts/ab_bench.sh base= unfused=ISDQN_FUSED_DENSE_ADAM=0
B="python bench.py --steps 300 --warmup 20 --capacity 20000 --cpu-seconds 1"
for spec in "$@"; do
  name="${spec%%=*}"; envs="${spec#*=}"
  env $envs timeout 200 $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -n 5 gpurun_out/ab_$name.err
done
python - "$@" <<'PY'
import json, sys
for spec in sys.argv[1:]:
    n = spec.split("=")[0]
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        ks = {k: v["ms"] for k, v in d["step_kernels_ms"].items() if "adam" in k or "reduce" in k or "dense_wgrad" in k}
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/step", round(d["ms_per_step"] * 1e3, 1), d["roofline"]["kernel"],
              round(d["roofline"]["frac"] or 0, 3), ks, "act_us", round(d["acting_us_per_action"], 1))
    except Exception as e:
        print(n, "ERR", e)
PY
python - "$@" <<'PY'
import json, sys
for spec in sys.argv[1:]:
    n = spec.split("=")[0]
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        ig = d.get("in_graph_us")
        if ig:
            print(n, "in-graph sum", round(sum(t for _, t in ig), 1), " ".join("%s:%.1f" % (k.replace("tc_", "")[:10], t) for k, t in ig))
    except Exception as e:
        print(n, "ERR", e)
PY
