"""Per-kernel times of one learner step (library launch marks, serialised) for an architecture / compute dtype:
    python scripts/step_profile.py cnn float32"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from isdqn_b200 import _lib
from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element, make_agent

arch = sys.argv[1] if len(sys.argv) > 1 else "cnn"
dt = sys.argv[2] if len(sys.argv) > 2 else "float32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
cfg = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch=arch)
el = batch_as_element(L.make_batch(1, B, cfg["obs_dim"], 9, arch))
agent = make_agent(1, **cfg, compute_dtype=dt, use_cuda_graph=False)
for _ in range(3):
    agent.learn_on_batch(agent.params, agent.optimizer_state, el)
torch.cuda.synchronize()
prof = _lib.profile(lambda: agent.learn_on_batch(agent.params, agent.optimizer_state, el))
print(arch, dt, "batch", B, "sum of launches ms", sum(t for _, t in prof))
for n, t in prof:
    print(f"{n:28s} {t * 1e3:9.1f} us")
