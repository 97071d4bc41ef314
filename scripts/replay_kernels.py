"""Runs the replay-side kernels once at their throughput shapes, for `ncu --set full` captures (profiles/r02_replay_ncu.md):
gather_stack4_u8 at 65,536 samples out of a 200 k-frame ring, sample_prioritized at 65,536 draws from a 1 M-key tree
(depth 21), sumtree_set_ops on a queue of add / add-then-evict ops, the batch-32 sumtree_set.

    ncu --set full --clock-control none -k regex:'gather_stack4|sample_prioritized|sumtree_set|uniform_' \
        -o gpurun_out/r02_replay python scripts/replay_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from isdqn_b200.sample_collection.replay_buffer import ReplayBuffer
from isdqn_b200.sample_collection.samplers import PrioritizedSamplingDistribution, UniformSamplingDistribution


def main():
    cap = int(os.environ.get("ISDQN_REPLAY_CAP", 200_000))
    rb = ReplayBuffer(UniformSamplingDistribution(0), 32, cap, stack_size=4, update_horizon=1, gamma=0.99,
                      frame_capacity=cap + cap // 8 + 64)
    bench.fill_replay(rb, 1000, cap + 1000)
    n_big = 65536
    n_keys = 1_000_000
    ps = PrioritizedSamplingDistribution(7, n_keys)
    rng = np.random.default_rng(7)
    pv = np.abs(rng.standard_normal(n_keys)) + 1e-3
    for k0 in range(0, n_keys, 65536):
        k1 = min(k0 + 65536, n_keys)
        ps._add_remove_run(k0, k1 - k0, 0, k1 - k0, pv[k0:k1].tolist())
    ps._sum_tree.flush()
    rb.sample_device(n_big)
    ps.sample_device(n_big, n_keys + 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()  # (ncu --profile-from-start off: the fills above are not captured)
    for _ in range(2):
        rb.sample_device(n_big)
    ps._add_remove_run(n_keys, 4096, 0, 0, (np.abs(rng.standard_normal(4096)) + 1e-3).tolist())  # add-then-evict ops
    ps._sum_tree.flush()
    for _ in range(2):
        ps.sample_device(n_big, n_keys + 1)
    ps._sum_tree.query(rng.random(n_big) * float(ps._sum_tree.root) * 0.999)
    keys = np.arange(5000, 5000 + 32 * 97, 97, dtype=np.int32)
    ps.update(keys, np.abs(rng.standard_normal(32)) + 1e-3)
    ps._sum_tree.flush()
    ps.update_device(torch.from_numpy(keys).cuda(), torch.from_numpy(np.abs(rng.standard_normal(32)) + 1e-3).cuda())
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ps.check_status()
    print("replay kernels done")


if __name__ == "__main__":
    main()
