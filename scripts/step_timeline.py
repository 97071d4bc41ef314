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
t = buf.cpu().numpy()
n = int(t[0])
per = n // n_rep
ts = t[1 : 1 + per * n_rep].reshape(n_rep, per).astype(np.float64)
print(f"{n} kernel starts, {per} per replay; names from profile: {len(names)}")
d = np.diff(np.concatenate([ts, np.roll(ts[:, :1], -1, axis=0)], axis=1), axis=1)[:-1]  # last column: to the next replay's first
med = np.median(d, axis=0) / 1e3
tot = 0.0
for i in range(per):
    nm = names[i] if i < len(names) else "?"
    print(f"{i:3d} {nm:22s} {med[i]:8.2f} us")
    tot += med[i]
print(f"sum {tot:.1f} us")
