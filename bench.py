#!/usr/bin/env python
"""bench.py — iS-DQN K=9 learner updates/s (+ replay samples/s) on B200.

    python bench.py --gpus 1 --steps 200 --warmup 20              # this repo (CUDA path)
    python bench.py --impl reference --steps 20 --warmup 3         # the reference's CPU path (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # N independent agents, one per GPU (weak scaling)

A "step" is one learner update of BASELINE.json configs[1]: draw 32 transitions from the replay buffer (uniform,
PCG64-exact), gather their 84x84x4 uint8 stacks out of the HBM frame ring, forward the K=9 Nature-CNN+LayerNorm
network on s and s', iterated TD targets + loss, backward, Adam.

  value  updates/s with every input resident in HBM (device sampler -> gather -> CUDA-graph learner step)
  e2e    updates/s through the reference-facing call `agent.learn_on_batch(params, opt_state, host_batch)`:
         pinned host batch -> H2D, step, D2H of the K losses — all inside the timed region
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FEATURES = [32, 64, 64, 512]
K_HEADS, N_ACTIONS, BATCH = 9, 9, 32
OBS = (84, 84, 4)
LR, ADAM_EPS, GAMMA = 6.25e-5, 1.5e-4, 0.99  # launch_job/atari/launch.sh, experiments/atari/isdqn.py:46
FLOP_PER_TRANSITION = 121_167_872


def synthetic_stream(seed: int, n: int, chunk: int = 4096):
    """Atari-shaped synthetic transitions: 90 % zero pixels, rewards in {-1,0,1}, geometric episodes."""
    rng = np.random.default_rng(seed)
    done = 0
    while done < n:
        m = min(chunk, n - done)
        frames = rng.integers(0, 256, (m, 84, 84), dtype=np.uint8)
        frames *= rng.random((m, 84, 84), dtype=np.float32) >= 0.9
        actions = rng.integers(0, N_ACTIONS, m)
        rewards = rng.integers(-1, 2, m).astype(np.float64)
        terminal = rng.random(m) < 1e-3
        for i in range(m):
            yield frames[i], int(actions[i]), float(rewards[i]), bool(terminal[i])
        done += m


def synthetic_chunks(seed: int, n: int, chunk: int = 4096):
    """The same stream as synthetic_stream, as arrays: (frames [m,84,84] u8, actions, rewards f64, terminal bool)."""
    rng = np.random.default_rng(seed)
    done = 0
    while done < n:
        m = min(chunk, n - done)
        frames = rng.integers(0, 256, (m, 84, 84), dtype=np.uint8)
        frames *= rng.random((m, 84, 84), dtype=np.float32) >= 0.9
        actions = rng.integers(0, N_ACTIONS, m)
        rewards = rng.integers(-1, 2, m).astype(np.float64)
        terminal = rng.random(m) < 1e-3
        yield frames, actions, rewards, terminal
        done += m


def fill_replay(rb, seed: int, n: int, priorities=None):
    """Fills `rb` with n synthetic transitions through the batched add path; returns the seconds spent inside
    `add_batch` (generating the synthetic frames is not the replay buffer's work)."""
    dt = 0.0
    for frames, a, r, d in synthetic_chunks(seed, n):
        t0 = time.perf_counter()
        rb.add_batch(frames, a, r, d, d, priorities=priorities)
        dt += time.perf_counter() - t0
    return dt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi needs ~100 ms to
    deliver its first sample and the timed region of a short run is a few milliseconds, so the sampler is started BEFORE
    the warm-up, every sample is stamped on arrival, and `summary()` keeps the samples that fall between `begin()` and
    `end()`.  A region that still caught none is followed by a probe: the caller re-runs the same steps untimed for
    ~0.3 s between `begin(probe=True)` / `end()`, and the record says so."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.windows, self.probe = [], False

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def begin(self, probe: bool = False):
        self._t0 = time.perf_counter()
        self.probe = self.probe or probe

    def end(self):
        self.windows.append((self._t0, time.perf_counter() + 0.03))  # (a sample is stamped when it arrives: one period late)

    def in_window(self) -> int:
        return sum(1 for t, _ in self.rows if any(a <= t <= b for a, b in self.windows))

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()

    def summary(self):
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if self.windows and not any(a <= t <= b for a, b in self.windows):
                continue
            try:
                sm.append(float(r[0]))
                smax = max(smax, float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
               "samples": len(sm)}
        if self.probe:
            out["note"] = ("the timed region was shorter than nvidia-smi's sampling period: sampled while the same steps were "
                           "re-run untimed right after it")
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_steps(n_steps: int, warmup: int, capacity: int = 20_000, time_budget_s: float = 1e9):
    """The reference's CPU learner path, restated (oracle port: JAX is not installable): ReplayOracle.sample(32) +
    PyTorch-CPU learn_on_batch on all host cores.  Returns (updates/s, steps done, cores, sample description)."""
    import torch

    from oracle import learner_oracle as L
    from oracle.replay_oracle import ReplayOracle
    from oracle.samplers_oracle import UniformSamplingOracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rb = ReplayOracle(UniformSamplingOracle(0), BATCH, capacity, 4, 1, GAMMA)
    for obs, a, r, d in synthetic_stream(0, capacity + 200):
        rb.add(obs, a, r, d, d)
    p = L.init_params(0, "cnn", OBS, FEATURES, (1 + K_HEADS) * N_ACTIONS, True, torch.float32)
    mu, nu, count = L.zeros_like_params(p), L.zeros_like_params(p), 0

    def step():
        nonlocal count
        b = rb.sample()
        batch = tuple(torch.from_numpy(np.asarray(x)) for x in b)
        count, losses, _, _, _ = L.learn_on_batch(p, mu, nu, count, batch, "cnn", True, K_HEADS, N_ACTIONS, GAMMA, 1, LR, ADAM_EPS)
        return losses

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    while done < n_steps and (time.perf_counter() - t0) < time_budget_s:
        step()
        done += 1
    dt = time.perf_counter() - t0
    desc = (f"{done} updates of batch {BATCH} (oracle port: numpy replay sample from a {capacity}-element buffer + "
            f"torch-CPU fp32 learn_on_batch, {cores} threads)")
    return done / dt, done, cores, desc, dt


def cpu_replay_baseline(capacity: int = 100_000, reps: int = 200):
    """the replay half of the metric on the host cores — the oracle restatements of rb.sample(32),
    SumTree.query(32), SumTree.set(32) and the prioritized sample(32) (single Python thread: the reference's replay side is
    single-threaded NumPy by construction).  Small frames would flatter rb.sample, so it runs at Atari shapes on a
    bounded buffer; the tree runs at the full 1 M capacity (depth 21)."""
    from oracle.replay_oracle import ReplayOracle
    from oracle.samplers_oracle import PrioritizedSamplingOracle, UniformSamplingOracle
    from oracle.sum_tree_oracle import SumTreeOracle

    out = {"cores": 1, "kind": "port", "unit": "samples/s"}
    n_el = min(capacity, 20_000)
    rb = ReplayOracle(UniformSamplingOracle(0), BATCH, n_el, 4, 1, GAMMA)
    for obs, a, r, d in synthetic_stream(0, n_el + 200):
        rb.add(obs, a, r, d, d)
    for _ in range(5):
        rb.sample()
    t0 = time.perf_counter()
    for _ in range(reps):
        rb.sample()
    out["rb_sample32_samples_per_s"] = BATCH * reps / (time.perf_counter() - t0)
    out["rb_sample32_sample"] = f"{reps} x ReplayOracle.sample(32), Atari shapes, {n_el}-element buffer (stacked copies: 56 KB each)"
    rng = np.random.default_rng(0)
    tree = SumTreeOracle(1_000_000)
    for lo in range(0, 1_000_000, 8192):  # distinct ascending leaves: one big set equals the single adds
        hi = min(lo + 8192, 1_000_000)
        tree.set(np.arange(lo, hi, dtype=np.int32), np.abs(rng.standard_normal(hi - lo)) + 1e-3)
    root = tree.root
    t0 = time.perf_counter()
    for _ in range(reps):
        tree.query(rng.random(BATCH) * root)
    out["sumtree_query32_samples_per_s"] = BATCH * reps / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for _ in range(reps):
        tree.set(rng.integers(0, 1_000_000, BATCH).astype(np.int32), np.abs(rng.standard_normal(BATCH)) + 1e-3)
    dt_set = time.perf_counter() - t0
    out["sumtree_set32_leaves_per_s"] = BATCH * reps / dt_set
    out["sumtree_set32_us"] = dt_set / reps * 1e6
    ps = PrioritizedSamplingOracle(1, 1_000_000)
    ps.tree = tree
    ps.index_to_key = list(range(1_000_000))
    t0 = time.perf_counter()
    for _ in range(reps):
        ps.sample(BATCH)
    out["prioritized_sample32_samples_per_s"] = BATCH * reps / (time.perf_counter() - t0)
    out["tree"] = "SumTreeOracle(1,000,000): depth 21, numpy float64"
    return out


REFERENCE_ARM_CAPACITY = 20_000  # elements the CPU arm's buffer holds (the oracle stores two stacked copies per element)


def run_reference(args, rank, world):
    if rank != 0:
        return
    ups, done, cores, desc, dt = cpu_reference_steps(args.steps, args.warmup, capacity=REFERENCE_ARM_CAPACITY)
    line = {
        "impl": "reference", "metric": "iS-DQN K=9 learner updates/sec", "value": ups, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(done, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the workload is OUR arm's (same network, batch, K, uniform replay); the CPU arm holds a bounded buffer — the oracle
        # keeps two stacked uint8 copies per element like the reference (56 KB each: 1 M elements would need 56 GB of
        # host memory) — and says so: `replay_capacity` is what it ran
        "config": dict(workload_config(REFERENCE_ARM_CAPACITY),
                       replay_capacity_note=f"bounded CPU sample: {REFERENCE_ARM_CAPACITY} of the CUDA arm's {args.capacity} elements; "
                                            "the sampled batch shape and the learner step are identical"),
        "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port", "sample": desc,
                         "replay": cpu_replay_baseline()},
        "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(capacity):
    return {
        "workload": "iS-DQN K=9 Nature-CNN+LayerNorm, Asterix-shaped synthetic uint8 84x84x4, batch 32, uniform replay "
                    "(BASELINE.json configs[1])",
        "batch": BATCH, "K": K_HEADS, "A": N_ACTIONS, "features": FEATURES, "replay_capacity": capacity,
        "parallelism": "independent agents, one per GPU, no communication (SURVEY §8e mode 1)",
        "l2": "inputs larger than L2: batches are drawn uniformly from the frame ring in HBM (7.06 GB at 1 M frames); "
              "parameters/Adam state (65 MB) stay L2-resident between steps as in real training",
    }


# ------------------------------------------------------------------------------------------------ CUDA arm
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the step
# (profiles/r02_b32_step_ncu.md, row 13; cold caches: the fp32 reads all come from DRAM, most writes are still in L2 at the end)
NCU_TRAFFIC_BYTES = {"adam": 67_726_848, "dense_wgrad_adam": 51_260_160}


DENSE0 = 7744 * 512  # the hidden Dense kernel (11*11*64 -> 512)


def kernel_work(name: str, P: int, fused: bool = False):
    """Algorithmic work of one launch for the roofline: ('hbm', bytes) or ('tensor', flops).
    fused: the hidden Dense kernel is updated by dense_wgrad_adam (the adam launch then covers the other leaves)."""
    B = BATCH
    if name == "adam":  # read p, g, mu, nu; write p, mu, nu (fp32) + the bf16 shadow of p
        return "hbm", 30 * (P - DENSE0 if fused else P)
    if name == "dense_wgrad_adam":  # read p, mu, nu; write p, mu, nu + bf16 shadow; the gradient is recomputed in registers
        return "hbm", 26 * DENSE0 + 2 * B * (7744 + 512)
    if name == "gather_stack4_u8":
        return "hbm", 91_728 * B
    if name == "dense_wgrad_gemm":
        return "hbm", 4 * 7744 * 512 + 4 * B * (7744 + 512)  # dW written once + operands read once
    if name == "dense_dgrad_gemm":
        return "hbm", 4 * 7744 * 512 + 4 * B * (7744 + 512)
    if name == "dense_fwd_gemm":
        return "hbm", 4 * 7744 * 512 + 4 * 2 * B * (7744 + 512)
    if name == "cast_params_bf16":
        return "hbm", 6 * P
    if name == "tc_dense_wgrad":
        return "hbm", 4 * 7744 * 512 + 2 * B * (7744 + 512)
    if name in ("tc_dense_dgrad", "tc_dense_fwd"):
        return "hbm", 2 * 7744 * 512 + 2 * 2 * B * (7744 + 512)
    if name in ("tc_conv_fwd", "tc_conv_fwd_tma"):  # three launches: torso forward on 2B images
        return "tensor", 2 * B * 24_076_288
    if name == "tc_conv_wgrad":
        return "tensor", B * 24_076_288
    if name in ("tc_conv_dgrad", "tc_conv_dgrad_tma"):
        return "tensor", B * (24_076_288 - 2 * 441 * 256 * 32)
    return None, 0


def measure_agent(agent, rb, steps, warmup, barrier, stream, local_rank):
    """updates/s of `agent` on `rb`: device-resident loop and the end-to-end loop (host batch -> H2D -> step -> D2H of the
    losses).  Returns (ms, ms_e2e, h2d bytes, d2h bytes, clocks summary); runs on `stream`."""
    import torch

    step_no = [0]

    def step():
        step_no[0] += 1
        agent.update_online_params(step_no[0], rb)

    clk = ClockSampler(local_rank).__enter__()  # (started before the warm-up: see ClockSampler)
    for _ in range(warmup):
        step()
    pool = [rb.sample() for _ in range(8)]
    h2d = sum(np.asarray(x).nbytes for x in pool[0])

    # 1-deep software pipeline, as a training loop runs it: step i+1 (host copy into pinned staging, H2D on the copy
    # engine, graph launch) is enqueued before step i's losses are read back; every step's H2D and D2H happen inside
    # the timed region.
    def e2e_loop(n):
        pending = None
        for i in range(n):
            agent.learn_on_batch(agent.params, agent.optimizer_state, pool[i % 8])
            handle = agent.losses_to_host_async()
            if pending is not None:
                pending.get()
            pending = handle
        return pending.get()

    e2e_loop(max(3, warmup // 2))
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk.begin()  # one window over both timed regions
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    e0.record()
    losses_host = e2e_loop(steps)
    e1.record()
    barrier()
    clk.end()
    if clk.in_window() == 0:  # too short for a sample: the same steps again, untimed, while nvidia-smi samples
        clk.begin(probe=True)
        t_probe = time.perf_counter()
        while time.perf_counter() - t_probe < 0.3:
            for _ in range(50):
                step()
            torch.cuda.synchronize()
        clk.end()
    clk.__exit__()
    return ev0.elapsed_time(ev1), e0.elapsed_time(e1), int(h2d), int(losses_host.size * 4), clk.summary()


def run_ours(args, rank, world, local_rank):
    import torch

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from isdqn_b200 import _lib
    from isdqn_b200.networks.isdqn import iSDQN
    from isdqn_b200.sample_collection.replay_buffer import ReplayBuffer, ReplayElement, TransitionElement
    from isdqn_b200.sample_collection.samplers import UniformSamplingDistribution

    cap = args.capacity
    rb = ReplayBuffer(UniformSamplingDistribution(rank), BATCH, cap, stack_size=4, update_horizon=1, gamma=GAMMA,
                      clipping=lambda x: np.clip(x, -1, 1), frame_capacity=cap + cap // 8 + 64, pinned_ring=16)
    n_fill = cap + max(cap // 10, 64)
    t_fill = fill_replay(rb, 1000 + rank, n_fill)
    # the per-transition path (what the reference's training loop calls) on a bounded sample, for the comparison
    n_one = 2000
    one = list(synthetic_stream(5000 + rank, n_one))
    t_one = time.perf_counter()
    for obs, a, r, d in one:
        rb.add(TransitionElement(obs, a, r, d, d))
    t_one = time.perf_counter() - t_one
    del one
    agent = iSDQN(rank, OBS, N_ACTIONS, K_HEADS, FEATURES, True, False, "cnn", LR, GAMMA, 1, 1, 8000, adam_eps=ADAM_EPS,
                  compute_dtype="bfloat16" if args.dtype == "bf16" else "float32")
    P = agent.network.n_params
    stream = torch.cuda.Stream()
    peaks = measured_peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        ms, ms_e2e, h2d, d2h, clocks = measure_agent(agent, rb, args.steps, args.warmup, barrier, stream, local_rank)
        # ---- the reference's own arithmetic (fp32, 1e-5 parity path) on the same buffer: a second agent, fewer steps
        fp32 = None
        if args.dtype == "bf16":
            agent32 = iSDQN(rank, OBS, N_ACTIONS, K_HEADS, FEATURES, True, False, "cnn", LR, GAMMA, 1, 1, 8000,
                            adam_eps=ADAM_EPS, compute_dtype="float32")
            n32 = max(20, args.steps // 4)
            ms32, ms32_e2e, _, _, _ = measure_agent(agent32, rb, n32, max(3, args.warmup // 4), barrier, stream, local_rank)
            fp32 = {"steps": n32, "ms": ms32, "ms_e2e": ms32_e2e}
            del agent32
        # ---- architecture_type "impala" (launch_job/atari/launch_time.sh runs both torsos with these features): fp32 path
        impala = None
        if args.dtype == "bf16" and rank == 0 and not args.no_impala:
            impala = {"features": list(FEATURES), "unit": "updates/s",
                      "scope": "rank 0 only; bf16: all convolutions and the hidden Dense layer on the tcgen05 tile engine, fp32 "
                               "residual stream / LayerNorm / head / Adam (2e-2 parity); f32: CUDA-core kernels (1e-5 parity)"}
            for name, cd in (("bf16", "bfloat16"), ("f32", "float32")):
                agent_i = iSDQN(rank, OBS, N_ACTIONS, K_HEADS, FEATURES, True, False, "impala", LR, GAMMA, 1, 1, 8000,
                                adam_eps=ADAM_EPS, compute_dtype=cd)
                n_i = max(10, args.steps // (4 if name == "bf16" else 10))
                ms_i, ms_i_e2e, _, _, _ = measure_agent(agent_i, rb, n_i, 3, lambda: torch.cuda.synchronize(), stream, local_rank)
                impala["params"] = agent_i.network.n_params
                impala[name] = {"steps": n_i, "value": n_i / (ms_i / 1e3), "ms_per_step": ms_i / n_i, "e2e": n_i / (ms_i_e2e / 1e3)}
                obs_i = np.random.default_rng(5).integers(0, 256, OBS, dtype=np.uint8)
                for i in range(10):
                    agent_i.best_action(agent_i.params, obs_i, i).item()
                t_i = time.perf_counter()
                for i in range(50):
                    agent_i.best_action(agent_i.params, obs_i, i).item()
                impala[name]["acting_us_per_action"] = (time.perf_counter() - t_i) / 50 * 1e6
                del agent_i
        # ---- replay throughput shape: 2048 batches of 32 per launch (sampler + gather only)
        n_big = 2048 * BATCH
        for _ in range(3):
            rb.sample_device(n_big)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        r0.record()
        for _ in range(reps):
            rb.sample_device(n_big)
        r1.record()
        torch.cuda.synchronize()
        ms_replay = r0.elapsed_time(r1) / reps
        prof_replay = _lib.profile(lambda: rb.sample_device(n_big))
        gather_ms = sum(t for n, t in prof_replay if n == "gather_stack4_u8")
        # ---- prioritized replay at the throughput shape: 65,536 draws per launch through the sum tree,
        # and the batch-32 priority update; 262,144 live keys (the Python-side fill of the key maps is what bounds this)
        prio = None
        if rank == 0:
            from isdqn_b200.sample_collection.samplers import PrioritizedSamplingDistribution

            n_keys = cap  # BASELINE configs[2]: 1 M keys, a tree of depth 21
            ps = PrioritizedSamplingDistribution(7, n_keys)
            prio_rng = np.random.default_rng(7)
            pv = np.abs(prio_rng.standard_normal(n_keys)) + 1e-3
            t_pfill = time.perf_counter()
            for k0 in range(0, n_keys, 65536):  # the add / add-then-evict call sequence of ReplayBuffer.add, batched
                k1 = min(k0 + 65536, n_keys)
                ps._add_remove_run(k0, k1 - k0, 0, k1 - k0, pv[k0:k1].tolist())
            n_evict = max(n_keys // 10, 64)
            pv2 = np.abs(prio_rng.standard_normal(n_evict)) + 1e-3
            ps._add_remove_run(n_keys, n_evict, 0, 0, pv2.tolist())  # every add evicts the oldest key (swap-remove, tagged set)
            ps._sum_tree.flush()
            torch.cuda.synchronize()
            t_pfill = time.perf_counter() - t_pfill
            for _ in range(3):
                ps.sample_device(n_big, n_keys + 1)
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(reps):
                ps.sample_device(n_big, n_keys + 1)
            p1.record()
            torch.cuda.synchronize()
            ms_prio = p0.elapsed_time(p1) / reps
            upd_keys = np.arange(n_evict, n_evict + 32 * 97, 97, dtype=np.int32)
            for _ in range(3):
                ps.update(upd_keys, np.abs(prio_rng.standard_normal(32)) + 1e-3)
                ps._sum_tree.flush()
            torch.cuda.synchronize()
            p0.record()
            for _ in range(50):
                ps.update(upd_keys, np.abs(prio_rng.standard_normal(32)) + 1e-3)
                ps._sum_tree.flush()
            p1.record()
            torch.cuda.synchronize()
            depth = ps._sum_tree._depth
            bytes_per_sample = (depth - 1) * 8 + 12
            prio = {"keys": n_keys, "tree_depth": depth, "samples_per_s": n_big / (ms_prio / 1e3), "launch_samples": n_big,
                    "ms_per_launch": ms_prio, "algorithmic_GBps": bytes_per_sample * n_big / (ms_prio / 1e3) / 1e9,
                    "update32_us": p0.elapsed_time(p1) / 50 * 1e3,
                    "fill": {"adds": n_keys + n_evict, "evictions": n_evict, "seconds": t_pfill,
                             "us_per_add": t_pfill / (n_keys + n_evict) * 1e6},
                    "bound": "L2 latency: the 16.8 MB heap is L2-resident, a draw is 20 dependent 8-byte reads; HBM is not the "
                             "limiter (profiles/r02_replay_ncu.md)"}
            # device-resident update: keys and priorities already on the GPU (the prioritized training driver's path)
            d_keys = torch.from_numpy(upd_keys).cuda()
            d_pr = torch.from_numpy(np.abs(prio_rng.standard_normal(32)) + 1e-3).cuda()
            for _ in range(3):
                ps.update_device(d_keys, d_pr)
            torch.cuda.synchronize()
            p0.record()
            for _ in range(50):
                ps.update_device(d_keys, d_pr)
            p1.record()
            torch.cuda.synchronize()
            ps.check_status()
            prio["update32_device_us"] = p0.elapsed_time(p1) / 50 * 1e3
            del ps
            # BASELINE configs[2] as a training loop: prioritized buffer (capacity = the uniform one's, inserts at the maximum
            # priority) -> draw with probabilities -> importance-weighted step -> |TD| written back as priorities, all on
            # the device, one graph replay per update (iSDQN.update_online_params with prioritized_beta set)
            if not args.no_impala:
                rbp = ReplayBuffer(PrioritizedSamplingDistribution(rank, cap), BATCH, cap, stack_size=4, update_horizon=1,
                                   gamma=GAMMA, clipping=lambda x: np.clip(x, -1, 1), frame_capacity=cap + cap // 8 + 64)
                t_lfill = fill_replay(rbp, 3000 + rank, n_fill, priorities="max")
                agent_p = iSDQN(rank, OBS, N_ACTIONS, K_HEADS, FEATURES, True, False, "cnn", LR, GAMMA, 1, 1, 8000,
                                adam_eps=ADAM_EPS, compute_dtype="bfloat16")
                agent_p.prioritized_beta = 0.4
                n_p = max(50, args.steps // 2)
                for i in range(10):
                    agent_p.update_online_params(i + 1, rbp)
                torch.cuda.synchronize()
                p0.record()
                for i in range(n_p):
                    agent_p.update_online_params(i + 1, rbp)
                p1.record()
                torch.cuda.synchronize()
                rbp._sampling_distribution.check_status()
                ms_p = p0.elapsed_time(p1)
                prio["learner"] = {"config": "BASELINE configs[2]: prioritized replay, capacity %d, K=9 cnn+LN, batch 32, beta 0.4" % cap,
                                   "steps": n_p, "value": n_p / (ms_p / 1e3), "unit": "updates/s", "ms_per_step": ms_p / n_p,
                                   "fill_us_per_add": t_lfill / n_fill * 1e6,
                                   "losses_finite": bool(np.isfinite(np.asarray(agent_p.losses_to_host_async().get())).all())}
                del agent_p, rbp
                torch.cuda.empty_cache()
        # ---- acting path: one greedy action for one host observation stack, read back with .item()
        act_state = np.random.default_rng(3).integers(0, 256, OBS, dtype=np.uint8)
        for i in range(20):
            agent.best_action(agent.params, act_state, i).item()
        t_act = time.perf_counter()
        for i in range(300):
            agent.best_action(agent.params, act_state, i).item()
        acting_us = (time.perf_counter() - t_act) / 300 * 1e6
        t_act = time.perf_counter()
        for i in range(300):
            int(agent.best_action_of_head(agent.params, act_state, i % K_HEADS))
        acting_head_us = (time.perf_counter() - t_act) / 300 * 1e6
        act_ctx = agent._ctx[1]["act"]
        act_fused = isinstance(act_ctx.get("fused"), torch.Tensor)
        acting = {"us_per_action": acting_us, "us_per_action_head_given": acting_head_us,
                  "path": ("single kernel, observation and actions through mapped pinned memory (isdqn_act_mapped)"
                           if act_fused and "h_flag" in act_ctx else
                           "single kernel (isdqn_act_host)" if act_fused else "layer chain (CUDA graph)")}
        if act_fused:
            lib_ = _lib.load()
            c1 = agent._ctx[1]
            prof_act = _lib.profile(lambda: lib_.isdqn_act(agent.network._net, agent.params.flat.data_ptr(), c1["state"].data_ptr(),
                                                           act_ctx["q"].data_ptr(), act_ctx["d_arg"].data_ptr(),
                                                           act_ctx["fused"].data_ptr(), act_ctx["fused"].numel(), stream.cuda_stream))
            acting["kernel_us"] = prof_act[0][1] * 1e3
        # batched acting: one forward for N environments, one synchronisation
        for n_env in (8, 64):
            states = np.random.default_rng(4).integers(0, 256, (n_env,) + OBS, dtype=np.uint8)
            keys = list(range(n_env))
            for _ in range(5):
                agent.best_actions(agent.params, states, keys)
            t_act = time.perf_counter()
            for _ in range(50):
                agent.best_actions(agent.params, states, keys)
            dt_act = (time.perf_counter() - t_act) / 50
            acting[f"best_actions_{n_env}_us_per_call"] = dt_act * 1e6
            acting[f"best_actions_{n_env}_us_per_action"] = dt_act * 1e6 / n_env
        # ---- per-kernel profile of one step (direct launches behind a spin kernel: no launch gaps)
        agent._use_graph = False
        prof = _lib.profile(lambda: agent.update_online_params(1, rb))
        agent._use_graph = True
        # ---- the same step INSIDE the graph replay: device-side kernel-start timeline (isdqn_trace_set); interval k =
        # start of kernel k -> start of kernel k+1 (duration + launch gap), median over 30 replays
        graph_names = [n for n, _ in prof if n not in ("sample_uniform", "gather_stack4_u8", "cast_params_bf16")]
        in_graph_us = None
        ctx = agent._context(BATCH)
        if ctx.get("graph") is not None:
            lib = _lib.load()
            tbuf = torch.zeros(4001, dtype=torch.int64, device="cuda")
            n_rep = 30
            torch.cuda.synchronize()
            _lib.check(lib.isdqn_trace_set(tbuf.data_ptr()), "isdqn_trace_set")
            for _ in range(n_rep):
                lib.isdqn_graph_launch(ctx["graph"], stream.cuda_stream)
            torch.cuda.synchronize()
            _lib.check(lib.isdqn_trace_set(None), "isdqn_trace_set")
            tr = tbuf.cpu().numpy().astype(np.uint64)
            ent = tr[1 : 1 + int(tr[0])]
            tg = (ent >> np.uint64(56)).astype(np.int64)
            tm = np.sort((ent & np.uint64(0x00FFFFFFFFFFFFFF)).astype(np.float64)[tg == 0])
            per = len(tm) // n_rep
            if per == len(graph_names) and per > 0:
                iv = np.diff(tm)[: per * (n_rep - 1)].reshape(n_rep - 1, per)
                in_graph_us = [(graph_names[k], float(np.median(iv[:, k]) / 1e3)) for k in range(per)]
    per_kernel = {}
    for name, t in prof:
        per_kernel.setdefault(name, [0, 0.0])
        per_kernel[name][0] += 1
        per_kernel[name][1] += t
    launches = len(prof)
    step_prof_ms = sum(t for _, t in prof)
    # dominant kernel = the longest single launch (three different convolutions share the name tc_conv_fwd)
    dominant = max(per_kernel.items(), key=lambda kv: kv[1][1] / kv[1][0])
    dom_name, (dom_count, dom_ms) = dominant
    fused = "dense_wgrad_adam" in per_kernel
    kind, work = kernel_work(dom_name, P, fused)

    t_max = torch.tensor([ms, ms_e2e, ms_replay] + ([fp32["ms"], fp32["ms_e2e"]] if fp32 else []), dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    t_host = [float(x) for x in t_max.cpu()]
    ms, ms_e2e, ms_replay = t_host[:3]
    if fp32:
        fp32 = {"dtype": "f32", "parity": "1e-5 (the reference's own arithmetic; CUDA-core FFMA path)", "steps": fp32["steps"],
                "value": world * fp32["steps"] / (t_host[3] / 1e3), "ms_per_step": t_host[3] / fp32["steps"],
                "e2e": world * fp32["steps"] / (t_host[4] / 1e3), "unit": "updates/s", "replay_capacity": cap}
    # release the agents-mode buffers before the large-batch measurements
    del rb, agent
    torch.cuda.empty_cache()
    # ---- BASELINE configs[4]: the data-parallel learner (global batch 4096, widths x1 and x2), strong scaling over the
    # same N GPUs; at N = 1 it is the single-device baseline the driver's efficiency is computed from
    dp = None
    if not args.no_dp:
        dp = {}
        for width in (1, 2):
            dp[f"w{width}"] = dp_measure(args, rank, world, local_rank, 4096, width, steps=max(10, min(args.steps // 10, 30)),
                                         warmup=5, dist=dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    value = world * args.steps / (ms / 1e3)
    e2e = world * args.steps / (ms_e2e / 1e3)
    roofline = None
    if kind == "tensor":
        ach = (work / 1e12) / (dom_ms / 1e3)  # `work` covers all launches of that name in one step
        roofline = {"kernel": dom_name, "bound": "tensor", "achieved": ach, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tflops_burst"], "traffic": None, "peak_src": peaks["src"],
                    "share_of_step": dom_ms / step_prof_ms, "launches_per_step": dom_count}
    elif kind == "hbm":
        ach = (work / 1e9) / ((dom_ms / dom_count) / 1e3)
        roofline = {"kernel": dom_name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": NCU_TRAFFIC_BYTES.get(dom_name), "peak_src": peaks["src"],
                    "share_of_step": dom_ms / step_prof_ms, "algorithmic_bytes": work}
    else:
        roofline = {"kernel": dom_name, "bound": "tensor", "achieved": None, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
                    "frac": None, "traffic": None, "peak_src": peaks["src"], "share_of_step": dom_ms / step_prof_ms}
    # every kernel of the step against its own bound (events, same measurement as `roofline`)
    roofline_kernels = []
    for name, (cnt, tot_ms) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
        k2, w2 = kernel_work(name, P, fused)
        if k2 == "tensor":
            a2 = (w2 / 1e12) / (tot_ms / 1e3)
            roofline_kernels.append({"kernel": name, "launches": cnt, "ms": round(tot_ms, 5), "bound": "tensor",
                                     "achieved": a2, "unit": "TFLOP/s", "frac": a2 / peaks["tflops_burst"]})
        elif k2 == "hbm":
            a2 = (w2 / 1e9) / ((tot_ms / cnt) / 1e3)
            roofline_kernels.append({"kernel": name, "launches": cnt, "ms": round(tot_ms, 5), "bound": "hbm",
                                     "achieved": a2, "unit": "GB/s", "frac": a2 / peaks["hbm_gbs"]})
    gather_gbs = 91_728 * n_big / (gather_ms / 1e3) / 1e9 if gather_ms > 0 else None
    # host-CPU baseline (rank 0 only: the other ranks have returned above), a bounded sample of the same workload
    ups, done, cores, desc, _ = cpu_reference_steps(10**9, 2, capacity=REFERENCE_ARM_CAPACITY, time_budget_s=args.cpu_seconds)
    cpu = {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port", "sample": desc,
           "replay": cpu_replay_baseline() if world == 1 else None}
    torso = None
    if dp and dp.get("w1"):
        torso = dict(dp["w1"]["torso_fwd"], batch=4096, width=1,
                     tensor_pipe_active_ncu="see profiles/ (ncu sm__pipe_tensor_op_hmma_cycles_active, not measurable live)")
    line = {
        "metric": "iS-DQN K=9 learner updates/sec", "value": value, "unit": "updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(cap),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps,
                "pipeline": "host numpy batches as rb.sample() returns them with pinned_ring=16 (views of a pinned block in "
                            "the packed batch layout) -> one H2D per step on the copy engine (2 device slots) -> step; the "
                            "losses of step i are read on the host after step i+1 is enqueued"},
        "gpu_launches": launches * args.steps,
        "gpu_launches_per_step": launches,
        "roofline": roofline,
        "roofline_kernels": roofline_kernels,
        "in_graph_us": in_graph_us,
        "cpu_baseline": cpu,
        "learner_tflops": world * args.steps * BATCH * FLOP_PER_TRANSITION / (ms / 1e3) / 1e12,
        "replay": {"samples_per_s": world * n_big / (ms_replay / 1e3), "launch_samples": n_big,
                   "gather_GBps": gather_gbs, "gather_frac_of_hbm": (gather_gbs / peaks["hbm_gbs"]) if gather_gbs else None,
                   "kernels_ms": {n: t for n, t in prof_replay}},
        "replay_prioritized": prio,
        "acting_us_per_action": acting_us,
        "acting": acting,
        "step_kernels_ms": {k: {"launches": v[0], "ms": round(v[1], 5)} for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])},
        "fill": {"adds": n_fill, "seconds": t_fill, "us_per_add": t_fill / n_fill * 1e6, "path": "ReplayBuffer.add_batch (4096 per call)",
                 "per_transition_add_us": t_one / n_one * 1e6},
        "fp32": fp32,
        "impala": impala,
        "torso": torso,
        "dp": dp,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def dp_measure(args, rank, world, local_rank, Bg, width, steps, warmup, dist=None):
    """One configuration of the data-parallel learner: global batch Bg split over `world` ranks (same draws on every rank,
    replicated storage), NCCL gradient all-reduce inside the step.  Returns the record (identical on every rank)."""
    import torch

    from isdqn_b200 import _lib
    from isdqn_b200.distributed import init_data_parallel, shard_bounds
    from isdqn_b200.networks.isdqn import iSDQN
    from isdqn_b200.sample_collection.replay_buffer import ReplayBuffer
    from isdqn_b200.sample_collection.samplers import UniformSamplingDistribution

    feats = [f * width for f in FEATURES]
    lo, hi = shard_bounds(Bg, rank, world)
    Bl = hi - lo
    cap = min(args.capacity, 50_000)
    rb = ReplayBuffer(UniformSamplingDistribution(0), Bg, cap, stack_size=4, update_horizon=1, gamma=GAMMA,
                      frame_capacity=cap + cap // 8 + 64)  # replicated storage, same seed => same draws on every rank
    fill_replay(rb, 1000, cap + 64)
    agent = iSDQN(0, OBS, N_ACTIONS, K_HEADS, feats, True, False, "cnn", LR, GAMMA, 1, 1, 8000, adam_eps=ADAM_EPS,
                  compute_dtype="bfloat16" if args.dtype == "bf16" else "float32")
    if world > 1:
        init_data_parallel(agent)
    stream = torch.cuda.Stream()
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        bufs = agent.batch_buffers(Bl)

        def step():
            _, _, d_slot = rb._sampling_distribution.sample_device(Bg, rb._slots)
            batch = rb._gather_slots_device(d_slot[lo:hi].contiguous(), out=bufs)
            agent.learn_on_batch(agent.params, agent.optimizer_state, batch)

        clk = ClockSampler(local_rank).__enter__()
        for _ in range(warmup):
            step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk.begin()
        ev0.record()
        for _ in range(steps):
            step()
        ev1.record()
        barrier()
        clk.end()
        ms = ev0.elapsed_time(ev1)
        # (the step contains a collective: whether to probe, and for how many steps, is decided by ALL ranks together)
        n_probe = max(1, int(0.3 / max(ms / steps / 1e3, 1e-6))) if clk.in_window() == 0 else 0
        if world > 1:
            t_n = torch.tensor([n_probe], dtype=torch.int64, device="cuda")
            dist.all_reduce(t_n, op=dist.ReduceOp.MAX)
            n_probe = int(t_n.item())
        if n_probe > 0:
            clk.begin(probe=True)
            for _ in range(n_probe):
                step()
            torch.cuda.synchronize()
            clk.end()
        clk.__exit__()
        use_graph = agent._use_graph
        agent._use_graph = False
        prof = _lib.profile(step)
        agent._use_graph = use_graph
    per_kernel = {}
    for name, t in prof:
        per_kernel.setdefault(name, [0, 0.0])
        per_kernel[name][0] += 1
        per_kernel[name][1] += t
    t_max = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms = float(t_max.item())
    P = agent.network.n_params
    flops = Bg * flops_per_transition(feats)
    torso = sum(v[1] for k, v in per_kernel.items() if k in ("tc_conv_fwd", "tc_conv_fwd_tma", "conv_fwd"))
    torso_flops = 2 * Bl * torso_fwd_flops(feats)
    ar = sum(v[1] for k, v in per_kernel.items() if k.startswith("nccl_allreduce"))
    rec = {
        "value": steps / (ms / 1e3), "unit": "updates/s", "scaling": "strong", "global_batch": Bg, "local_batch": Bl, "width": width,
        "features": feats, "params": P, "steps": steps, "ms_per_step": ms / steps,
        "transitions_per_s": Bg * steps / (ms / 1e3), "learner_tflops": flops * steps / (ms / 1e3) / 1e12,
        "allreduce": {"bytes": 4 * P, "ms_serial": ar, "note": "duration of the all-reduce launches in a serialised, event-timed "
                      "replay of one step (rank 0); inside the timed steps part of it overlaps the backward pass"} if world > 1 else None,
        "torso_fwd": {"ms": torso, "tflops": torso_flops / (torso / 1e3) / 1e12 if torso > 0 else None,
                      "frac_of_sustained": (torso_flops / (torso / 1e3) / 1e12 / peaks["tflops_sustained"]) if torso > 0 else None,
                      "peak_tflops_sustained": peaks["tflops_sustained"]},
        "clocks": clk.summary(),
        "step_kernels_ms": {k: {"launches": v[0], "ms": round(v[1], 4)} for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])},
    }
    del rb, agent
    torch.cuda.empty_cache()
    return rec


def run_dp(args, rank, world, local_rank):
    """BASELINE configs[4] on its own (`--mode dp`): large-batch (default 4096, CNN x2) data-parallel learner."""
    import torch

    torch.cuda.set_device(local_rank)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    Bg, width = args.batch or 4096, args.width or 2
    rec = dp_measure(args, rank, world, local_rank, Bg, width, args.steps, args.warmup, dist=dist)
    if rank == 0:
        line = {
            "metric": "iS-DQN K=9 learner updates/sec (data parallel)", "value": rec["value"], "unit": "updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"large-batch iS-DQN K=9 data parallel, global batch {Bg}, CNN width x{width} "
                                   "(BASELINE.json configs[4])", "global_batch": Bg, "features": rec["features"], "params": rec["params"],
                       "parallelism": f"dp{world}: NCCL all-reduce of the {4 * rec['params'] / 1e6:.1f} MB fp32 gradient"},
            "clocks": rec["clocks"], "transitions_per_s": rec["transitions_per_s"], "learner_tflops": rec["learner_tflops"],
            "torso_fwd": rec["torso_fwd"], "allreduce": rec["allreduce"], "step_kernels_ms": rec["step_kernels_ms"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def torso_fwd_flops(feats):
    """2 * MACs of the three convs for one image (SAME padding: 84 -> 21 -> 11 -> 11)."""
    return 2 * (441 * 256 * feats[0] + 121 * 16 * feats[0] * feats[1] + 121 * 9 * feats[1] * feats[2])


def flops_per_transition(feats):
    """fwd on 2 images + wgrad + dgrad (no dgrad for the first conv), SURVEY §8(d) accounting."""
    c0, c1, c2 = 441 * 256 * feats[0], 121 * 16 * feats[0] * feats[1], 121 * 9 * feats[1] * feats[2]
    d0, d1 = 121 * feats[2] * feats[3], feats[3] * (1 + K_HEADS) * N_ACTIONS
    fwd = 2 * (c0 + c1 + c2 + d0 + d1)
    return 2 * fwd + fwd + (fwd - 2 * c0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--capacity", type=int, default=int(os.environ.get("ISDQN_BENCH_CAPACITY", 1_000_000)))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--mode", default="agents", choices=["agents", "dp"],
                    help="agents: one independent agent per GPU (weak scaling, BASELINE configs[1]/[3]); dp: ONE learner, "
                         "global batch split over the GPUs, NCCL gradient all-reduce (strong scaling, configs[4])")
    ap.add_argument("--batch", type=int, default=None, help="dp mode: global batch (default 4096)")
    ap.add_argument("--width", type=int, default=None, choices=[1, 2, 4], help="dp mode: CNN width multiplier (default 2)")
    ap.add_argument("--no-dp", action="store_true", help="agents mode: skip the data-parallel sub-records (configs[4])")
    ap.add_argument("--no-impala", action="store_true", help="skip the impala sub-record (profiling runs)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"],
                    help="bf16: tcgen05 tensor-core path (fp32 accumulate/master weights); f32: CUDA-core fp32 parity path")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.mode == "dp":
        if args.warmup < 3:
            args.warmup = 3
        run_dp(args, rank, world, local_rank)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
