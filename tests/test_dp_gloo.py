"""CPU, world_size 2, gloo: host-side logic of the data-parallel mode (is-dqn_b200/distributed.py).  The compute on
each rank is the oracle (no GPU here); what is checked is the sharding rule, the 1/B_global scaling and that
all-reducing the per-rank gradient shares reproduces the single-process gradient of the concatenated batch — the
oracle SURVEY.md §8e names for the DP mode."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from isdqn_b200.distributed import loss_scale, shard_batch, shard_bounds
        from isdqn_b200.sample_collection.replay_buffer import ReplayElement
        from oracle import learner_oracle as L

        torch.set_num_threads(2)
        K, A, B = 3, 4, 8
        feats = [8, 8, 8, 16]
        p = L.init_params(0, "cnn", (20, 20, 2), feats, (1 + K) * A, True)
        L.randomize_small_leaves(p, 1)
        full = L.make_batch(5, B, (20, 20, 2), A, "cnn")  # the SAME global draw on every rank
        el = ReplayElement(*[x.numpy() for x in full])
        mine = shard_batch(el, rank, world)
        lo, hi = shard_bounds(B, rank, world)
        assert mine.action.shape[0] == B // world and np.array_equal(mine.state, el.state[lo:hi])
        local = tuple(torch.from_numpy(np.asarray(x)) for x in mine)
        _, losses, grads, _, _ = L.learn_on_batch(L.clone_params(p), L.zeros_like_params(p), L.zeros_like_params(p), 0, local,
                                                  "cnn", True, K, A, 0.99, 1, 0.0, 1.0)
        scale = loss_scale(B // world, B)
        flat = torch.cat([g.reshape(-1) for m in grads.values() for g in m.values()]) * scale
        share = losses * scale
        dist.all_reduce(flat)
        dist.all_reduce(share)
        _, ref_losses, ref_grads, _, _ = L.learn_on_batch(L.clone_params(p), L.zeros_like_params(p), L.zeros_like_params(p), 0,
                                                          full, "cnn", True, K, A, 0.99, 1, 0.0, 1.0)
        ref_flat = torch.cat([g.reshape(-1) for m in ref_grads.values() for g in m.values()])
        assert torch.allclose(flat, ref_flat, rtol=1e-10, atol=1e-12), float((flat - ref_flat).abs().max())
        assert torch.allclose(share, ref_losses, rtol=1e-12)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_dp_gradient_shares_sum_to_the_single_process_gradient(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_bounds_cover_the_batch():
    from isdqn_b200.distributed import shard_bounds

    for world in (1, 2, 4, 8):
        cuts = [shard_bounds(4096, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == 4096
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)
