"""One impala learner step (batch 32, bf16 mode, direct launches) between cudaProfilerStart / Stop:
    ncu --set full --clock-control none --profile-from-start off -o /tmp/impala python scripts/impala_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element, make_agent

cfg = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch="impala")
agent = make_agent(1, **cfg, compute_dtype=sys.argv[1] if len(sys.argv) > 1 else "bfloat16", use_cuda_graph=False)
el = batch_as_element(L.make_batch(1, 32, cfg["obs_dim"], 9, "impala"))
for _ in range(3):
    agent.learn_on_batch(agent.params, agent.optimizer_state, el)
torch.cuda.synchronize()
torch.cuda.profiler.start()
agent.learn_on_batch(agent.params, agent.optimizer_state, el)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
