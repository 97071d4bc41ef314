"""Where an action's time goes: host pieces timed with perf_counter, the kernel with the library's
per-launch marks.  Run on a GPU box: python scripts/acting_profile.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isdqn_b200 import _lib  # noqa: E402
from isdqn_b200 import _lib  # noqa: E402
from isdqn_b200.networks.isdqn import iSDQN, _raw_key  # noqa: E402

OBS = (84, 84, 4)


def per_call_us(fn, n=2000, warm=50):
    for i in range(warm):
        fn(i)
    t = time.perf_counter()
    for i in range(n):
        fn(i)
    return (time.perf_counter() - t) / n * 1e6


def main():
    agent = iSDQN(0, OBS, 9, 9, [32, 64, 64, 512], True, False, "cnn", 6.25e-5, 0.99, 1, 1, 8000, adam_eps=1.5e-4,
                  compute_dtype=os.environ.get("DTYPE", "bfloat16"))
    obs = np.random.default_rng(3).integers(0, 256, OBS, dtype=np.uint8)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        out = {}
        out["best_action (key -> head draw -> action)"] = per_call_us(lambda i: agent.best_action(agent.params, obs, i).item())
        out["best_action_of_head"] = per_call_us(lambda i: int(agent.best_action_of_head(agent.params, obs, i % 9)))
        out["head draw alone (default_rng(SeedSequence).integers)"] = per_call_us(
            lambda i: int(_lib.load().isdqn_threefry_randint(*_raw_key(i), 0, 9)))
        a = agent._ctx[1]["act"]
        out["pinned copy of the observation alone"] = per_call_us(lambda i: a["h_obs_np"].__setitem__(Ellipsis, obs.reshape(-1)))
        lib = _lib.load()
        net = agent.network
        ctx = agent._ctx[1]
        if a["fused"] is not False:
            args = (net._net, agent.params.flat.data_ptr(), a["h_obs"].data_ptr(), ctx["state"].data_ptr(), a["nbytes"],
                    a["q"].data_ptr(), a["d_arg"].data_ptr(), a["h_arg"].data_ptr(), a["fused"].data_ptr(), a["fused"].numel(),
                    stream.cuda_stream, a["ev"])
            out["isdqn_act_host alone (H2D + kernel + D2H + sync)"] = per_call_us(lambda i: lib.isdqn_act_host(*args))
            prof = _lib.profile(lambda: lib.isdqn_act(net._net, agent.params.flat.data_ptr(), ctx["state"].data_ptr(),
                                                       a["q"].data_ptr(), a["d_arg"].data_ptr(), a["fused"].data_ptr(),
                                                       a["fused"].numel(), stream.cuda_stream))
            out["act_forward kernel (device, us)"] = [round(ms * 1e3, 2) for _, ms in prof]
            # warm L2 vs after a learner-sized sweep of other memory
            junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            ts = []
            for _ in range(5):
                junk.zero_()
                prof = _lib.profile(lambda: lib.isdqn_act(net._net, agent.params.flat.data_ptr(), ctx["state"].data_ptr(),
                                                           a["q"].data_ptr(), a["d_arg"].data_ptr(), a["fused"].data_ptr(),
                                                           a["fused"].numel(), stream.cuda_stream))
                ts.append(round(prof[0][1] * 1e3, 2))
            out["act_forward kernel, L2 flushed (us)"] = ts
        if a["fused"] is not False:
            # phase marks of CTA 0 inside the kernel (globaltimer): tag 0 = start, 10 + 2l = layer l computed, 11 + 2l = past
            # the barrier after layer l
            tbuf = torch.zeros(4001, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            _lib.check(lib.isdqn_trace_set(tbuf.data_ptr()), "isdqn_trace_set")
            reps = 20
            for _ in range(reps):
                lib.isdqn_act(net._net, agent.params.flat.data_ptr(), ctx["state"].data_ptr(), a["q"].data_ptr(), a["d_arg"].data_ptr(),
                              a["fused"].data_ptr(), a["fused"].numel(), stream.cuda_stream)
                torch.cuda.synchronize()
            _lib.check(lib.isdqn_trace_set(None), "isdqn_trace_set")
            tr = tbuf.cpu().numpy().astype(np.uint64)
            ent = tr[1 : 1 + int(tr[0])]
            tags = (ent >> np.uint64(56)).astype(np.int64)
            tm = (ent & np.uint64(0x00FFFFFFFFFFFFFF)).astype(np.float64)
            per = len(ent) // reps
            rel = (tm.reshape(reps, per) - tm.reshape(reps, per)[:, :1]) / 1e3
            out["phase marks (tag: us since kernel start, median)"] = [f"{int(t)}:{np.median(rel[:, k]):.1f}" for k, t in enumerate(tags[:per])]
    for k, v in out.items():
        print(f"{k:60s} {v if isinstance(v, list) else round(v, 2)}")


if __name__ == "__main__":
    main()
