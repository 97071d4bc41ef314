"""Where does a batch-32 step spend its time: GPU (graph replay) or host (Python between launches)?

    python scripts/host_overhead.py            # on a GPU box

Prints the device time of back-to-back graph replays, the wall/device time of the full `update_online_params` loop, and
a cProfile of the host side of that loop.
"""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from isdqn_b200 import _lib  # noqa: E402
from isdqn_b200.networks.isdqn import iSDQN  # noqa: E402
from isdqn_b200.sample_collection.replay_buffer import ReplayBuffer, TransitionElement  # noqa: E402
from isdqn_b200.sample_collection.samplers import UniformSamplingDistribution  # noqa: E402

cap = 20000
rb = ReplayBuffer(UniformSamplingDistribution(0), 32, cap, stack_size=4, update_horizon=1, gamma=0.99,
                  clipping=lambda x: np.clip(x, -1, 1), frame_capacity=cap + cap // 8 + 64)
for obs, a, r, d in bench.synthetic_stream(1000, cap + 2000):
    rb.add(TransitionElement(obs, a, r, d, d))
agent = iSDQN(0, bench.OBS, bench.N_ACTIONS, bench.K_HEADS, bench.FEATURES, True, False, "cnn", bench.LR, bench.GAMMA, 1, 1, 8000,
              adam_eps=bench.ADAM_EPS, compute_dtype="bfloat16")
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    for i in range(20):
        agent.update_online_params(i + 1, rb)
    torch.cuda.synchronize()
    ctx = agent._context(32)
    lib = _lib.load()
    n = 500
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        lib.isdqn_graph_launch(ctx["graph"], stream.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    print(f"graph replay only        : {e0.elapsed_time(e1) / n * 1e3:8.1f} us/step (device)")

    e0.record()
    for _ in range(n):
        rb.sample_device(out=agent.batch_buffers(32))
    e1.record()
    torch.cuda.synchronize()
    print(f"sample_device only       : {e0.elapsed_time(e1) / n * 1e3:8.1f} us/step (device)")
    t0 = time.perf_counter()
    for _ in range(n):
        rb.sample_device(out=agent.batch_buffers(32))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"sample_device host time  : {(t1 - t0) / n * 1e6:8.1f} us/step (no sync)")

    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        agent.update_online_params(i + 1, rb)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"update_online_params     : {e0.elapsed_time(e1) / n * 1e3:8.1f} us/step (device), host enqueue {(t1 - t0) / n * 1e6:.1f} us/step, "
          f"drain {(t2 - t1) * 1e6:.0f} us")
    pr = cProfile.Profile()
    pr.enable()
    for i in range(n):
        agent.update_online_params(i + 1, rb)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
