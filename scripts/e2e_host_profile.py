import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from isdqn_b200.networks.isdqn import iSDQN
from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element
agent = iSDQN(0, bench.OBS, bench.N_ACTIONS, bench.K_HEADS, bench.FEATURES, True, False, "cnn", bench.LR, bench.GAMMA, 1, 1, 8000,
              adam_eps=bench.ADAM_EPS, compute_dtype="bfloat16")
pool = [batch_as_element(L.make_batch(i, 32, bench.OBS, bench.N_ACTIONS, "cnn")) for i in range(8)]
stream = torch.cuda.Stream()
def loop(n):
    pending = None
    for i in range(n):
        agent.learn_on_batch(agent.params, agent.optimizer_state, pool[i % 8])
        h = agent.losses_to_host_async()
        if pending is not None: pending.get()
        pending = h
    pending.get()
with torch.cuda.stream(stream):
    loop(50)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(500); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("e2e loop: %.1f us/step" % ((t1 - t0) / 500 * 1e6))
    pr = cProfile.Profile(); pr.enable(); loop(500); pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
