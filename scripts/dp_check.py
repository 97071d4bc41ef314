"""GPU, N ranks (torchrun): the data-parallel step against the single-device step on the concatenated batch
(the oracle SURVEY.md §8e names for this mode).  Prints one line per check; exits non-zero on mismatch.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from isdqn_b200.distributed import allreduce_losses, init_data_parallel, shard_batch
from isdqn_b200.networks.isdqn import iSDQN
from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # impala / float32: the first update agrees to 2e-7; from the second on a handful of the ~1.5 M ReLU units per sample sit
    # within the 1e-7 parameter difference of zero and take the other branch, which Adam's division by sqrt(v) + eps turns
    # into 2e-5 .. 5e-5 of the largest parameter (the same 1e-4 bar as tests/test_learner_gpu.py TOL_PARAM)
    for arch, dtype, tol in (("cnn", "float32", 2e-5), ("cnn", "bfloat16", 2e-2), ("impala", "float32", 1e-4),
                             ("impala", "bfloat16", 2e-2)):
        Bg = (32 if arch == "cnn" else 8) * world
        mk = lambda: iSDQN(7, (84, 84, 4), 9, 9, [32, 64, 64, 512], True, False, arch, 6.25e-5, 0.99, 1, 1, 10**9,
                           adam_eps=1.5e-4, compute_dtype=dtype)
        dp = mk()
        init_data_parallel(dp)
        single = mk()  # same seed => same initial parameters; trains on the full batch on every rank
        for step in range(3):
            full = batch_as_element(L.make_batch(100 + step, Bg, (84, 84, 4), 9, "cnn"))
            _, _, l_dp = dp.learn_on_batch(dp.params, dp.optimizer_state, shard_batch(full, rank, world))
            l_dp = allreduce_losses(l_dp)
            _, _, l_1 = single.learn_on_batch(single.params, single.optimizer_state, full)
            e_l = float((l_dp - l_1).abs().max() / l_1.abs().max())
            e_p = float((dp.params.flat - single.params.flat).abs().max() / single.params.flat.abs().max())
            good = e_l <= tol and e_p <= tol
            ok &= good
            if rank == 0:
                print(f"dp_check {arch} {dtype} world={world} step={step} losses rel err {e_l:.2e} params rel err {e_p:.2e} {'OK' if good else 'FAIL'}", flush=True)
        # parameters must be bit-identical across ranks
        ref = dp.params.flat.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, dp.params.flat))
        ok &= same
        if rank == 0:
            print(f"dp_check {arch} {dtype} replicas bit-identical: {same}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
