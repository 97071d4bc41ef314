"""Kernel-start timeline of one batch-32 learner step INSIDE the CUDA-graph replay (isdqn_trace_set).

    python scripts/step_timeline.py [n_replays]

Prints, for every kernel of the step, the median interval from its start to the start of the next kernel (= its
duration + the launch gap behind it).
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from isdqn_b200 import _lib  # noqa: E402
from isdqn_b200.networks.isdqn import iSDQN  # noqa: E402
from oracle import learner_oracle as L  # noqa: E402
from tests.learner_utils import batch_as_element  # noqa: E402

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 50
agent = iSDQN(0, bench.OBS, bench.N_ACTIONS, bench.K_HEADS, bench.FEATURES, True, False, "cnn", bench.LR, bench.GAMMA, 1, 1, 8000,
              adam_eps=bench.ADAM_EPS, compute_dtype="bfloat16")
el = batch_as_element(L.make_batch(0, 32, bench.OBS, bench.N_ACTIONS, "cnn"))
lib = _lib.load()
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    for _ in range(5):
        agent.learn_on_batch(agent.params, agent.optimizer_state, el)
    torch.cuda.synchronize()
    ctx = agent._context(32)
    # names from the per-kernel profile of a direct (non-graph) step
    agent._use_graph = False
    names = [n for n, _ in _lib.profile(lambda: agent.learn_on_batch(agent.params, agent.optimizer_state, el))]
    agent._use_graph = True
    names = [n for n in names if n not in ("cast_params_bf16",)]
    buf = torch.zeros(4001, dtype=torch.int64, device="cuda")
    _lib.check(lib.isdqn_trace_set(buf.data_ptr()), "trace_set")
    for _ in range(n_rep):
        lib.isdqn_graph_launch(ctx["graph"], stream.cuda_stream)
    torch.cuda.synchronize()
    _lib.check(lib.isdqn_trace_set(None), "trace_set")
t = buf.cpu().numpy().astype(np.uint64)
n = int(t[0])
ent = t[1 : 1 + n]
tags = (ent >> np.uint64(56)).astype(np.int64)
times = (ent & np.uint64(0x00FFFFFFFFFFFFFF)).astype(np.float64)
order = np.argsort(times, kind="stable")
tags, times = tags[order], times[order]
starts = np.flatnonzero(tags == 0)
per = len(starts) // n_rep
print(f"{n} entries, {len(starts)} kernel starts, {per} per replay; names from profile: {len(names)}")
# per kernel slot: interval to the next kernel start, and the phase marks (relative to the kernel start)
rows = {}
for si, s0 in enumerate(starts[:-1]):
    slot = si % per
    s1 = starts[si + 1]
    rec = rows.setdefault(slot, {"dt": [], "ph": {}})
    rec["dt"].append(times[s1] - times[s0])
    for e in range(s0 + 1, s1):
        rec["ph"].setdefault(int(tags[e]), []).append(times[e] - times[s0])
tot = 0.0
print("  # kernel                 start->next |  prologue  tables  1st-chunk  mma-issued  acc-ready  cta0-done   (us from kernel start)")
for slot in range(per):
    rec = rows[slot]
    nm = names[slot] if slot < len(names) else "?"
    med = np.median(rec["dt"]) / 1e3
    tot += med
    ph = "".join(f"{np.median(rec['ph'][k]) / 1e3:10.2f}" if k in rec["ph"] else "         -" for k in range(1, 7))
    print(f"{slot:3d} {nm:22s} {med:8.2f}    |{ph}")
print(f"sum {tot:.1f} us")
