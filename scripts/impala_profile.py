"""Per-kernel times of one impala learner step (library launch marks, serialised) and the graph-replayed step time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from isdqn_b200 import _lib
from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element, make_agent

cfg = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch="impala")
dt = sys.argv[1] if len(sys.argv) > 1 else "bfloat16"
el = batch_as_element(L.make_batch(1, 32, cfg["obs_dim"], 9, "impala"))
agent = make_agent(1, **cfg, compute_dtype=dt)
for _ in range(5):
    agent.learn_on_batch(agent.params, agent.optimizer_state, el)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    agent.learn_on_batch(agent.params, agent.optimizer_state, el)
torch.cuda.synchronize()
print(dt, "graph-replayed step", (time.perf_counter() - t0) / 50 * 1e3, "ms")
agent = make_agent(1, **cfg, compute_dtype=dt, use_cuda_graph=False)
for _ in range(3):
    agent.learn_on_batch(agent.params, agent.optimizer_state, el)
torch.cuda.synchronize()
prof = _lib.profile(lambda: agent.learn_on_batch(agent.params, agent.optimizer_state, el))
agg = {}
for n, t in prof:
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += t
print("sum of launches ms", sum(v[1] for v in agg.values()))
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:28s} x{v[0]:3d} {v[1] * 1e3:9.1f} us")
