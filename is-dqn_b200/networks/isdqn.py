"""iS-DQN agent on the device — drop-in for `slimdqn/networks/isdqn.py` (same constructor, attributes, methods).

One shared torso, `(1 + K) * A` outputs reshaped to `(batch, 1 + K, A)`: head 0 is the frozen target, heads 1..K are
online and head k regresses onto `r + gamma^n (1 - d) max_a Q_{k-1}(s')` (isdqn.py:34-43, 92-109).

`learn_on_batch` is ONE native call (`isdqn_learn_on_batch`: forward on s and s', fused K-head TD loss fwd+bwd,
backward, [NCCL all-reduce], Adam), captured once in a CUDA graph and replayed; params, Adam moments and the step
counter stay in HBM as flat float32 vectors.  The reference's functional signature is kept — the returned pytrees
are the (updated in place) inputs, i.e. the arguments are donated.
"""
from __future__ import annotations

import os

import numpy as np

from .. import _lib
from ..sample_collection.replay_buffer import ReplayElement
from .architectures.dqn import DQNNet, ParamTree


def _raw_key(key):
    """A JAX PRNG key as its two uint32 words.  uint32[2] arrays (what `jax.random.PRNGKey` / `split` produce) are taken
    as they are; a Python int seed is `jax.random.PRNGKey(seed)` with x64 disabled: (0, seed mod 2^32)."""
    if isinstance(key, (int, np.integer)):
        return 0, int(key) & 0xFFFFFFFF
    k = np.asarray(key)
    if k.dtype.kind in "iu" and k.size == 2:
        k = k.reshape(-1)
        return int(k[0]) & 0xFFFFFFFF, int(k[1]) & 0xFFFFFFFF
    # anything else (test harnesses hand over arbitrary seeds): fold the bytes into two words
    b = np.frombuffer(k.tobytes(), dtype=np.uint8)
    h = np.frombuffer(np.random.SeedSequence(b.astype(np.uint32) if b.size else [0]).generate_state(2).tobytes(), dtype=np.uint32)
    return int(h[0]), int(h[1])


def _new_event(lib):
    ev = _lib.C.c_void_p()
    _lib.check(lib.isdqn_event_create(ev), "isdqn_event_create")
    return ev


class HostLosses:
    """Handle on an in-flight device->host copy of one step's losses (`iSDQN.losses_to_host_async`)."""

    def __init__(self, lib, buf, event):
        self._lib, self._buf, self._event = lib, buf, event

    def get(self) -> np.ndarray:
        _lib.check(self._lib.isdqn_event_synchronize(self._event), "isdqn_event_synchronize")
        return self._buf.copy()


class OptState(dict):
    """`optax.adam` state: {"count": int32[1], "mu": ParamTree, "nu": ParamTree} (optax 0.2.4 ScaleByAdamState)."""


class iSDQN:
    def __init__(
        self,
        key,
        observation_dim,
        n_actions,
        n_bellman_iterations: int,
        features: list,
        layer_norm: bool,
        batch_norm: bool,
        architecture_type: str,
        learning_rate: float,
        gamma: float,
        update_horizon: int,
        data_to_update: int,
        target_update_frequency: int,
        adam_eps: float = 1e-8,
        use_cuda_graph: bool = True,
        compute_dtype: str = "float32",
    ):
        torch = _lib.require_cuda()
        self._torch = torch
        self._lib = _lib.load()
        self.n_bellman_iterations = int(n_bellman_iterations)
        self.n_actions = int(n_actions)
        features = [int(f) for f in features]
        self.last_idx_mlp = len(features) if architecture_type == "fc" else len(features) - 3
        self.network = DQNNet(
            features, architecture_type, (1 + self.n_bellman_iterations) * self.n_actions, layer_norm, batch_norm
        )
        self.network.configure(observation_dim, self.n_bellman_iterations, self.n_actions)

        # 1 + self.n_bellman_iterations = [\bar{Q_0}, Q_1, ..., Q_K]
        def apply(params, state):
            q_values, batch_stats = self.network.apply(params, state, mutable=["batch_stats"])
            return q_values.reshape((-1, 1 + self.n_bellman_iterations, self.n_actions)), batch_stats

        self.network.apply_fn = apply
        self.params = self.network.init(key)

        # optax.adam(learning_rate, eps=adam_eps): b1=0.9, b2=0.999, eps_root=0 (isdqn.py:46)
        self.learning_rate = float(learning_rate)
        self.adam_eps = float(adam_eps)
        self.adam_b1, self.adam_b2 = 0.9, 0.999
        self.optimizer_state = OptState(
            count=torch.zeros(1, dtype=torch.int32, device="cuda"),
            mu=self.network.new_tree(),
            nu=self.network.new_tree(),
        )
        self._grads = self.network.new_tree()

        self.gamma = gamma
        self.update_horizon = update_horizon
        self.data_to_update = data_to_update
        self.target_update_frequency = target_update_frequency
        self._d_cumulated = torch.zeros(self.n_bellman_iterations, dtype=torch.float64, device="cuda")

        self._use_graph = bool(use_cuda_graph)
        # "float32": CUDA-core FFMA path (1e-5 parity); "bfloat16": tcgen05 tensor-core path for the conv torso and
        # the hidden Dense layers, fp32 accumulation / master weights / Adam (2e-2 parity, north_star)
        if compute_dtype not in ("float32", "bfloat16"):
            raise ValueError(f"compute_dtype must be 'float32' or 'bfloat16', got {compute_dtype!r}")
        self.compute_dtype = compute_dtype
        self._ctx = {}  # batch size -> persistent buffers + captured graph
        self._nccl_comm = None
        self._dp_world = 1
        self._dp_rank = 0
        self._side_stream = None
        self._copy_stream = None
        self._loss_ring = None
        self._last_step = None
        # prioritized replay (set by the training script): beta of the importance weights (None = plain replay, the
        # reference's behaviour) and the constant added to |TD| before it becomes a priority
        self.prioritized_beta = None
        self.prioritized_eps = 1e-6

    def td_abs(self, B: int):
        """float32 CUDA tensor [K][B]: |TD| of every online head and sample of the latest step on a batch of B."""
        return self._context(B)["td_abs"]

    # ------------------------------------------------------------------------------------- loss bookkeeping
    @property
    def cumulated_losses(self) -> np.ndarray:
        """Host view of the per-head loss sums (isdqn.py:53,62).  Accumulated on the device: reading it is the only
        synchronisation, instead of one per update."""
        return self._d_cumulated.cpu().numpy()

    @cumulated_losses.setter
    def cumulated_losses(self, value) -> None:
        self._d_cumulated.copy_(self._torch.as_tensor(np.asarray(value, dtype=np.float64)))

    # ----------------------------------------------------------------------------------------- data parallel
    def enable_data_parallel(self, comm_handle: int, world_size: int, rank: int = 0) -> None:
        """Large-batch data-parallel mode (new functionality): every rank runs the step on its
        slice of the global batch; gradients are all-reduced over NCCL before the (replicated) Adam step."""
        self._nccl_comm = comm_handle
        self._dp_world = int(world_size)
        self._dp_rank = int(rank)
        self._ctx.clear()

    # ------------------------------------------------------------------------------------------- step context
    def _context(self, B: int):
        ctx = self._ctx.get(B)
        if ctx is not None:
            return ctx
        t = self._torch
        net = self.network
        ctx = {}
        if net.image_input:
            shape, dt = (B,) + tuple(net.observation_dim), t.uint8
        else:
            shape, dt = (B,) + tuple(net.observation_dim), t.float32
        # the five batch fields live in ONE device allocation (16-byte aligned sub-ranges) so that a staged host batch
        # reaches them with a single device-to-device copy
        item = 1 if dt == t.uint8 else 4
        o, offs = _lib.batch_pack_layout(B, int(np.prod(shape[1:])) * item)
        ctx["pack_bytes"], ctx["pack_offs"] = o, offs
        ctx["dev_pack"] = t.zeros(o, dtype=t.uint8, device="cuda")

        def views(pack):
            v = {}
            for name, vdt, vshape in (("state", dt, shape), ("next_state", dt, shape), ("action", t.int64, (B,)),
                                      ("reward", t.float64, (B,)), ("terminal", t.uint8, (B,))):
                a, nb = offs[name]
                v[name] = pack[a : a + nb].view(vdt).view(vshape)
            return v

        ctx["views"] = views
        ctx.update(views(ctx["dev_pack"]))
        ctx["losses"] = t.zeros(self.n_bellman_iterations, dtype=t.float32, device="cuda")
        # prioritized training driver: |TD| of every (online head, sample) of the latest step, and the importance
        # weights the loss applies when a prioritized batch is learned from (ones otherwise: never read then)
        ctx["td_abs"] = t.zeros((self.n_bellman_iterations, B), dtype=t.float32, device="cuda")
        ctx["is_weights"] = t.ones(B, dtype=t.float32, device="cuda")
        nbytes = self._lib.isdqn_learn_workspace_bytes(net._net, B)
        if nbytes < 0:
            raise _lib.IsdqnNativeError("isdqn_learn_workspace_bytes failed")
        ctx["ws"] = t.empty(max(nbytes, 16), dtype=t.uint8, device="cuda")
        ctx["ws_tc"] = None
        if self.compute_dtype == "bfloat16":
            nb = self._lib.isdqn_learn_workspace_tc_bytes(net._net, B)
            if nb <= 0:
                raise _lib.IsdqnNativeError(
                    "compute_dtype='bfloat16' needs a cnn with 32/64/128/256-channel convolutions and hidden Dense "
                    "widths that are multiples of 64, or an impala network with 32/64/128/256-channel stacks; use "
                    "compute_dtype='float32' for this network"
                )
            ctx["ws_tc"] = t.empty(nb, dtype=t.uint8, device="cuda")
        # double-buffered pinned + device staging for host batches (the reference's implicit device_put at the jit
        # boundary): allocated on first use by _stage_host_batch
        ctx["stage"] = None
        batch = _lib.Batch(
            ctx["state"].data_ptr(), ctx["next_state"].data_ptr(), ctx["action"].data_ptr(),
            ctx["reward"].data_ptr(), ctx["terminal"].data_ptr(),
        )
        ctx["batch"] = batch
        ctx["graph"] = None
        ctx["graph_key"] = None
        ctx["warm"] = 0
        self._ctx[B] = ctx
        return ctx

    def batch_buffers(self, B: int) -> ReplayElement:
        """The persistent device buffers the captured step reads: a replay buffer can gather straight into them
        (`rb.sample_device(out=agent.batch_buffers(B))`)."""
        c = self._context(B)
        return ReplayElement(c["state"], c["action"], c["reward"], c["next_state"], c["terminal"])

    def _ensure_shadow(self, params: ParamTree) -> None:
        if params.shadow is None:
            params.shadow = self._torch.empty(params.flat.numel(), dtype=self._torch.bfloat16, device="cuda")
            params.shadow_dirty = True

    def _refresh_shadow(self, params: ParamTree, stream) -> None:
        """bf16 shadow <- fp32 parameters (needed after the parameters were written outside the captured step)."""
        self._ensure_shadow(params)
        _lib.check(
            self._lib.isdqn_cast_f32_to_bf16(params.flat.data_ptr(), params.shadow.data_ptr(), params.flat.numel(), stream),
            "isdqn_cast_f32_to_bf16",
        )
        params.shadow_dirty = False

    def _train_struct(self, ctx, params: ParamTree, opt: OptState, B: int, refresh_shadow: bool = True) -> "_lib.Train":
        tr = _lib.Train()
        tr.gamma_n = float(self.gamma**self.update_horizon)
        tr.lr, tr.b1, tr.b2, tr.eps = self.learning_rate, self.adam_b1, self.adam_b2, self.adam_eps
        tr.batch = B
        tr.batch_global = B * self._dp_world
        tr.d_params = params.flat.data_ptr()
        tr.d_grads = self._grads.flat.data_ptr()
        tr.d_mu = opt["mu"].flat.data_ptr() if opt is not None else None
        tr.d_nu = opt["nu"].flat.data_ptr() if opt is not None else None
        tr.d_count = opt["count"].data_ptr() if opt is not None else None
        tr.d_losses = ctx["losses"].data_ptr()
        tr.d_td_abs = ctx["td_abs"].data_ptr()
        tr.d_is_weights = None
        tr.d_workspace = ctx["ws"].data_ptr()
        tr.workspace_bytes = ctx["ws"].numel()
        tr.nccl_comm = self._nccl_comm
        tr.compute_dtype = _lib.COMPUTE_BF16 if self.compute_dtype == "bfloat16" else _lib.COMPUTE_F32
        if ctx["ws_tc"] is not None:
            self._ensure_shadow(params)
            tr.d_workspace_tc = ctx["ws_tc"].data_ptr()
            tr.workspace_tc_bytes = ctx["ws_tc"].numel()
            tr.d_params_bf16 = params.shadow.data_ptr()
            # direct launches always rebuild the shadow (5 us); the captured graph relies on Adam keeping it current
            tr.refresh_shadow = 1 if refresh_shadow else 0
            if refresh_shadow:
                params.shadow_dirty = False
        return tr

    def _stage_host_batch(self, ctx, fields, stream, in_place_ok: bool = False):
        """Host numpy batch -> (pinned slot ->) device slot on the copy stream -> the batch buffers on the step stream.
        Two slots: the H2D copy of step i+1 runs on the copy engine while step i computes; the CPU only blocks when it
        is two steps ahead of the GPU.  A batch that already lives in a pinned block of the packed layout (what
        `ReplayBuffer.sample()` returns) is copied straight from there: no host-side copy at all."""
        t = self._torch
        lib = self._lib
        st = ctx["stage"]
        if st is None:
            st = ctx["stage"] = {
                "slot": 0,
                "host": [_lib.pinned_block(ctx["pack_bytes"]) for _ in range(2)],
                # (torch.empty: a zero-fill would be a kernel on the CURRENT stream, which the copy stream does not follow
                # — it could land after the first batch was staged and wipe it)
                "dev": [t.empty(ctx["pack_bytes"], dtype=t.uint8, device="cuda") for _ in range(2)],
                "h2d_done": [_new_event(lib) for _ in range(2)],
                "stage_free": [_new_event(lib) for _ in range(2)],
            }
            st["host_np"] = [{k: v.numpy() for k, v in ctx["views"](h).items()} for h in st["host"]]
            if self._copy_stream is None:
                self._copy_stream = t.cuda.Stream()
            # whatever the allocator did to these blocks on the current stream is ordered before both users
            cur = t.cuda.current_stream()
            self._copy_stream.wait_stream(cur)
            stream.wait_stream(cur)
        slot = st["slot"]
        st["slot"] = slot ^ 1
        arrays = [np.asarray(fields[k]) for k in _lib.BATCH_FIELDS]
        if self.network.image_input and (arrays[0].dtype != np.uint8 or arrays[1].dtype != np.uint8):
            raise TypeError("cnn / impala batches must be uint8 frames (float states: use loss_on_batch / apply)")
        h_src = _lib.pinned_pack_base(arrays, ctx["pack_offs"])
        if h_src is None:
            # pageable (or differently laid out) arrays: through this slot's pinned block, once its last copy has left
            _lib.check(lib.isdqn_event_synchronize(st["h2d_done"][slot]), "isdqn_event_synchronize")
            for name, arr in zip(_lib.BATCH_FIELDS, arrays):
                dst = st["host_np"][slot][name]
                dst[...] = arr.reshape(dst.shape)
            h_src = st["host"][slot].data_ptr()
        # the H2D copy lands in this slot's device block and the step reads it THERE (one captured graph per slot): no
        # device-to-device copy into the persistent batch buffers on the step's critical path
        in_place = in_place_ok and os.environ.get("ISDQN_STAGE_INPLACE", "1") != "0"
        _lib.check(
            lib.isdqn_stage_batch(h_src, st["dev"][slot].data_ptr(), None if in_place else ctx["dev_pack"].data_ptr(),
                                  ctx["pack_bytes"], self._copy_stream.cuda_stream, stream.cuda_stream, st["h2d_done"][slot],
                                  st["stage_free"][slot]),
            "isdqn_stage_batch",
        )
        if in_place:
            if "batch" not in st:
                st["batch"] = []
                for d in st["dev"]:
                    v = ctx["views"](d)
                    st["batch"].append(_lib.Batch(v["state"].data_ptr(), v["next_state"].data_ptr(), v["action"].data_ptr(),
                                                  v["reward"].data_ptr(), v["terminal"].data_ptr()))
            return slot
        return None

    def _load_batch(self, ctx, batch, stream, in_place_ok: bool = False):
        """Copies `batch` (host numpy, like `rb.sample()`; or CUDA tensors) into the persistent device buffers, ordered
        on `stream` (a torch.cuda.Stream).  Returns the staging slot the step has to read instead (host batches with the
        in-place staging), or None."""
        t = self._torch
        names = ("state", "action", "reward", "next_state", "terminal")
        fields = dict(zip(names, (batch.state, batch.action, batch.reward, batch.next_state, batch.is_terminal)))
        if not any(isinstance(f, t.Tensor) for f in fields.values()):
            return self._stage_host_batch(ctx, fields, stream, in_place_ok)
        with t.cuda.stream(stream):
            for name, f in fields.items():
                dst = ctx[name]
                if isinstance(f, t.Tensor):
                    if f.data_ptr() == dst.data_ptr():
                        continue
                    dst.copy_(f.reshape(dst.shape) if f.dtype == dst.dtype else f.reshape(dst.shape).to(dst.dtype), non_blocking=True)
                else:
                    dst.copy_(t.as_tensor(np.asarray(f)).reshape(dst.shape).to(dst.dtype), non_blocking=False)
        return None

    def losses_to_host_async(self) -> "HostLosses":
        """Enqueues the device->host copy of the latest step's K losses behind that step and returns a handle;
        `handle.get()` blocks on THAT copy only, so a training loop can read step i's losses while step i+1 runs
        (jax's asynchronous dispatch gives the reference the same overlap).  A handle stays valid for the next 8 calls."""
        t = self._torch
        lib = self._lib
        if self._loss_ring is None:
            buf = _lib.pinned_block(8 * 4 * self.n_bellman_iterations).view(t.float32).view(8, self.n_bellman_iterations)
            self._loss_ring = {"i": 0, "buf": buf, "np": buf.numpy(), "ev": [_new_event(lib) for _ in range(8)]}
        if self._last_step is None:
            raise RuntimeError("losses_to_host_async() follows a learn_on_batch() call")
        losses, stream = self._last_step
        r = self._loss_ring
        i = r["i"]
        r["i"] = (i + 1) % 8
        _lib.check(
            lib.isdqn_read_async(r["buf"][i].data_ptr(), losses.data_ptr(), 4 * self.n_bellman_iterations, stream.cuda_stream,
                                 r["ev"][i]),
            "isdqn_read_async",
        )
        return HostLosses(lib, r["np"][i], r["ev"][i])

    # ------------------------------------------------------------------------------------------------ update
    def update_online_params(self, step: int, replay_buffer):
        if step % self.data_to_update == 0:
            device_rb = hasattr(replay_buffer, "sample_device") and self.network.image_input
            sd = getattr(replay_buffer, "_sampling_distribution", None)
            if device_rb and self.prioritized_beta is not None and hasattr(sd, "update_device"):
                # prioritized training driver (new: the reference never wires its prioritized sampler to an agent,
                # SURVEY F10): draw with probabilities, learn with importance weights, write |TD| back — all on the device
                if (self._dp_world == 1 and self._use_graph and os.environ.get("ISDQN_GRAPH_SAMPLE", "1") != "0"
                        and self._learn_from_replay(replay_buffer, prioritized=True)):
                    return
                B = replay_buffer._batch_size
                batch_samples, d_keys, d_w = replay_buffer.sample_device(out=self.batch_buffers(B), beta=self.prioritized_beta)
                self.params, self.optimizer_state, losses = self.learn_on_batch(
                    self.params, self.optimizer_state, batch_samples, _accumulate=True, is_weights=d_w
                )
                replay_buffer.update_device(d_keys, self.td_abs(B), prio_rows=self.n_bellman_iterations,
                                            offset=self.prioritized_eps)
                return
            if (device_rb and self._dp_world == 1 and self._use_graph and os.environ.get("ISDQN_GRAPH_SAMPLE", "1") != "0"
                    and self._learn_from_replay(replay_buffer)):
                return
            if self._dp_world > 1:
                # data parallel: the replay buffer's batch is the GLOBAL batch — the same draw on every rank (replicated
                # storage, same sampler seed) — and this rank learns from its contiguous slice of it
                from ..distributed import shard_batch, shard_bounds

                Bg = replay_buffer._batch_size
                lo, hi = shard_bounds(Bg, self._dp_rank, self._dp_world)
                if device_rb:
                    _, _, d_slot = replay_buffer._sampling_distribution.sample_device(Bg, replay_buffer._slots)
                    batch_samples = replay_buffer._gather_slots_device(d_slot[lo:hi].contiguous(), out=self.batch_buffers(hi - lo))
                else:
                    batch_samples = shard_batch(replay_buffer.sample(), self._dp_rank, self._dp_world)
            elif device_rb:
                B = replay_buffer._batch_size
                batch_samples = replay_buffer.sample_device(out=self.batch_buffers(B))
            else:
                batch_samples = replay_buffer.sample()

            # (`self.cumulated_losses += losses`, isdqn.py:62, happens inside the step's loss kernel)
            self.params, self.optimizer_state, losses = self.learn_on_batch(
                self.params, self.optimizer_state, batch_samples, _accumulate=True
            )

    def update_target_params(self, step: int):
        if step % self.target_update_frequency == 0:
            # Window shift
            self.params = self.shift_params(self.params)

            if self._dp_world > 1:
                # every rank holds its share of the global means (loss sums use 1 / B_global): K floats, summed when logged
                import torch.distributed as dist

                if dist.is_available() and dist.is_initialized():
                    if dist.get_backend() == "nccl":
                        dist.all_reduce(self._d_cumulated)
                    else:
                        h = self._d_cumulated.cpu()
                        dist.all_reduce(h)
                        self._d_cumulated.copy_(h)
            cumulated = self.cumulated_losses
            logs = {
                "loss": np.mean(cumulated) / (self.target_update_frequency / self.data_to_update),
            }
            for idx_network in range(min(self.n_bellman_iterations, 5)):
                logs[f"networks/{idx_network}_loss"] = cumulated[idx_network] / (
                    self.target_update_frequency / self.data_to_update
                )
            self._d_cumulated.zero_()

            return True, logs

        return False, {}

    def _learn_from_replay(self, replay_buffer, prioritized: bool = False) -> bool:
        """`rb.sample()` + `learn_on_batch` as ONE graph replay: draw -> gather -> step, and for the prioritized driver
        draw with importance weights -> gather -> weighted step -> |TD| written back as priorities.  The host pushes what
        it has pending (frames, records, key-map patches, sum-tree ops) before the replay.  False: not applicable."""
        B = replay_buffer._batch_size
        ctx = self._context(B)
        name = "rb_plan_prio" if prioritized else "rb_plan"
        plan = ctx.get(name)
        if plan is None or plan[0] is not replay_buffer:
            if prioritized:
                made = (replay_buffer.capturable_prioritized_step(self.batch_buffers(B), ctx["is_weights"])
                        if hasattr(replay_buffer, "capturable_prioritized_step") else None)
            else:
                made = replay_buffer.capturable_sample(self.batch_buffers(B)) if hasattr(replay_buffer, "capturable_sample") else None
            if made is None:
                return False
            plan = ctx[name] = (replay_buffer,) + tuple(made)
        post = None
        if prioritized:
            _, prepare, enqueue, update, set_beta, token = plan
            td_abs, K, eps = ctx["td_abs"], self.n_bellman_iterations, float(self.prioritized_eps)
            post = lambda stream_ptr: update(stream_ptr, td_abs, K, eps)
        else:
            _, prepare, enqueue, token = plan
        cur = self._torch.cuda.current_stream()
        side = None
        if cur.cuda_stream == 0:  # the legacy default stream cannot be captured
            if self._side_stream is None:
                self._side_stream = self._torch.cuda.Stream()
            side = self._side_stream
            side.wait_stream(cur)
        run = side if side is not None else cur
        with self._torch.cuda.stream(run):
            prepare()
            if prioritized:
                set_beta(self.prioritized_beta)
        try:
            out = self._learn_on_stream(ctx, self.params, self.optimizer_state, B, run.cuda_stream, True, prioritized,
                                        pre=enqueue, pre_token=(token(), eps) if prioritized else token(), post=post)
            self._last_step = (out[2], run)
        finally:
            if side is not None:
                cur.wait_stream(side)
        return True

    def learn_on_batch(self, params: ParamTree, optimizer_state: OptState, batch_samples, _accumulate: bool = False,
                       is_weights=None):
        """isdqn.py:82-90.  Returns (params, optimizer_state, losses[K] float32 CUDA tensor); params / optimizer
        state are updated in place (donated) and returned.  is_weights (optional, [B]): importance weights of a
        prioritized batch — losses[k] = mean_b w_b td^2."""
        B = int(batch_samples.action.shape[0])
        ctx = self._context(B)
        cur = self._torch.cuda.current_stream()
        side = None
        if self._use_graph and cur.cuda_stream == 0:
            # the legacy default stream cannot be captured: run the step on a side stream ordered after it
            if self._side_stream is None:
                self._side_stream = self._torch.cuda.Stream()
            side = self._side_stream
            side.wait_stream(cur)
        run = side if side is not None else cur
        slot = self._load_batch(ctx, batch_samples, run, in_place_ok=True)
        if is_weights is not None:
            with self._torch.cuda.stream(run):
                w = is_weights if isinstance(is_weights, self._torch.Tensor) else self._torch.as_tensor(np.asarray(is_weights))
                ctx["is_weights"].copy_(w.reshape(B).to(self._torch.float32), non_blocking=True)
        try:
            out = self._learn_on_stream(ctx, params, optimizer_state, B, run.cuda_stream, _accumulate, is_weights is not None,
                                        batch_slot=slot)
            if slot is not None:  # the slot may be overwritten once this step has read it
                st = ctx["stage"]
                _lib.check(self._lib.isdqn_event_record(st["stage_free"][slot], run.cuda_stream), "isdqn_event_record")
            self._last_step = (out[2], run)
            return out
        finally:
            if side is not None:
                cur.wait_stream(side)

    def _learn_on_stream(self, ctx, params, optimizer_state, B, stream, accumulate=False, weighted=False, pre=None,
                         pre_token=None, batch_slot=None, post=None):
        """pre(stream) / post(stream): launches enqueued in front of / behind the step (the replay draw + gather; the
        priority write-back), captured with it.
        batch_slot: the step reads its batch from that host-batch staging slot instead of the persistent batch buffers."""
        batch = ctx["batch"] if batch_slot is None else ctx["stage"]["batch"][batch_slot]
        key = (params.flat.data_ptr(), optimizer_state["mu"].flat.data_ptr(), optimizer_state["nu"].flat.data_ptr(),
               optimizer_state["count"].data_ptr(), stream, bool(accumulate), bool(weighted), pre_token, batch_slot)
        # a few captured variants are kept side by side (host-batch step, replay-fed step, weighted step): a training loop
        # that alternates between them must not re-capture at every switch
        graphs = ctx.setdefault("graphs", {})
        if self._use_graph and key in graphs:
            if ctx["ws_tc"] is not None and params.shadow_dirty:
                self._refresh_shadow(params, stream)
            ctx["graph"], ctx["graph_key"] = graphs[key], key
            _lib.check(self._lib.isdqn_graph_launch(ctx["graph"], stream), "isdqn_graph_launch")
            return params, optimizer_state, ctx["losses"]
        tr = self._train_struct(ctx, params, optimizer_state, B)
        tr.d_cumulated = self._d_cumulated.data_ptr() if accumulate else None
        tr.d_is_weights = ctx["is_weights"].data_ptr() if weighted else None
        # (data parallel: the NCCL all-reduces are captured with the step — NCCL supports stream capture; ISDQN_DP_GRAPH=0
        # keeps direct launches)
        if self._use_graph and ctx["warm"] >= 1 and (self._nccl_comm is None or os.environ.get("ISDQN_DP_GRAPH", "1") != "0"):
            # capture this very step (it executes on replay, not during capture)
            if len(graphs) >= 6:  # (stale variants: parameters were re-allocated, another replay buffer, ...)
                for g in graphs.values():
                    self._lib.isdqn_graph_destroy(g)
                graphs.clear()
                ctx["graph"] = None
            if ctx["ws_tc"] is not None:
                self._refresh_shadow(params, stream)  # outside the capture
                tr.refresh_shadow = 0
            _lib.check(self._lib.isdqn_graph_begin(stream), "isdqn_graph_begin")
            err = None
            try:
                if pre is not None:
                    pre(stream)
            except _lib.IsdqnNativeError as e:  # (the capture must be closed whatever happened inside it)
                err = e
            rc = self._lib.isdqn_learn_on_batch(self.network._net, tr, batch, stream)
            try:
                if post is not None and err is None and rc == 0:
                    post(stream)
            except _lib.IsdqnNativeError as e:
                err = e
            exec_ = _lib.C.c_void_p()
            rc2 = self._lib.isdqn_graph_end(stream, exec_)
            if err is not None:
                raise err
            _lib.check(rc, "isdqn_learn_on_batch (capture)")
            _lib.check(rc2, "isdqn_graph_end")
            ctx["graph"], ctx["graph_key"] = exec_, key
            graphs[key] = exec_
            _lib.check(self._lib.isdqn_graph_launch(ctx["graph"], stream), "isdqn_graph_launch")
        else:
            if pre is not None:
                pre(stream)
            _lib.check(self._lib.isdqn_learn_on_batch(self.network._net, tr, batch, stream), "isdqn_learn_on_batch")
            if post is not None:
                post(stream)
            ctx["warm"] += 1
        return params, optimizer_state, ctx["losses"]

    def grad_on_batch(self, params: ParamTree, batch_samples):
        """Gradient of the loss without the update: (grads ParamTree, losses[K]).  Used by parity tests / DP."""
        B = int(batch_samples.action.shape[0])
        ctx = self._context(B)
        self._load_batch(ctx, batch_samples, self._torch.cuda.current_stream())
        tr = self._train_struct(ctx, params, None, B)
        _lib.check(self._lib.isdqn_grad_on_batch(self.network._net, tr, ctx["batch"], _lib.stream_ptr()), "isdqn_grad_on_batch")
        return self._grads, ctx["losses"]

    def loss_on_batch(self, params: ParamTree, samples):
        """isdqn.py:92-103: (sum_k mean_b td^2, (per-head means (K,), batch_stats)).  Also leaves the (2B, 1+K, A)
        Q-values of the call in `self.last_all_q_values`."""
        t = self._torch
        net = self.network
        state, _, is_float = net.prepare_input(samples.state)
        next_state, _, _ = net.prepare_input(samples.next_state)
        B = int(state.shape[0])
        if net.image_input and is_float:
            # float states (tests/utils.py generator): forward through the float entry point, then the loss kernel
            all_q, _ = net.apply_fn(params, t.cat((state, next_state)))
            all_q = all_q.reshape(2 * B, -1).contiguous()
            losses = t.empty(self.n_bellman_iterations, dtype=t.float32, device="cuda")
            # named references: the buffers must outlive the (asynchronous) kernel in allocator order
            d_action = self._as_dev(samples.action, t.int64)
            d_reward = self._as_dev(samples.reward, t.float64)
            d_terminal = self._as_dev(samples.is_terminal, t.uint8)
            _lib.check(
                self._lib.isdqn_heads_td_loss(
                    all_q.data_ptr(), d_action.data_ptr(), d_reward.data_ptr(), d_terminal.data_ptr(),
                    float(self.gamma**self.update_horizon), B, B,
                    self.n_bellman_iterations, self.n_actions, losses.data_ptr(), None, _lib.stream_ptr(),
                ),
                "isdqn_heads_td_loss",
            )
        else:
            ctx = self._context(B)
            self._load_batch(ctx, ReplayElement(state, samples.action, samples.reward, next_state, samples.is_terminal),
                             t.cuda.current_stream())
            tr = self._train_struct(ctx, params, None, B)
            all_q = t.empty((2 * B, net.final_feature), dtype=t.float32, device="cuda")
            _lib.check(
                self._lib.isdqn_loss_on_batch(net._net, tr, ctx["batch"], all_q.data_ptr(), _lib.stream_ptr()),
                "isdqn_loss_on_batch",
            )
            losses = ctx["losses"].clone()
        self.last_all_q_values = all_q.reshape(2 * B, 1 + self.n_bellman_iterations, self.n_actions)
        return losses.sum(), (losses, {})

    def _as_dev(self, x, dtype):
        t = self._torch
        if isinstance(x, t.Tensor):
            return x.to("cuda", dtype).contiguous()
        return t.as_tensor(np.asarray(x)).to("cuda", dtype).contiguous()

    def compute_target(self, sample, next_q_values):
        """isdqn.py:105-109: r + (1 - d) * gamma^n * max_a next_q (fp32, evaluated as r + (((1-d) gamma^n) max))."""
        t = self._torch
        nq = next_q_values if isinstance(next_q_values, t.Tensor) else t.as_tensor(np.asarray(next_q_values))
        nq = nq.to(t.float32)
        r = t.as_tensor(np.asarray(sample.reward.cpu() if isinstance(sample.reward, t.Tensor) else sample.reward), dtype=t.float32).to(nq.device)
        d = t.as_tensor(np.asarray(sample.is_terminal.cpu() if isinstance(sample.is_terminal, t.Tensor) else sample.is_terminal)).to(nq.device).to(t.int32)
        coef = (1 - d).to(t.float32) * t.tensor(self.gamma**self.update_horizon, dtype=t.float32, device=nq.device)
        return r + coef * nq.max(dim=-1).values

    def shift_params(self, params: ParamTree) -> ParamTree:
        """isdqn.py:111-125: \\bar{Q}_i <- Q_{i+1} on the last Dense layer only (Adam moments are not shifted)."""
        mod = params["params"][f"Dense_{self.last_idx_mlp}"]
        kernel, bias = mod["kernel"], mod["bias"]
        _lib.check(
            self._lib.isdqn_shift_heads(
                kernel.data_ptr(), bias.data_ptr(), int(kernel.shape[0]), self.n_bellman_iterations, self.n_actions,
                _lib.stream_ptr(),
            ),
            "isdqn_shift_heads",
        )
        params.shadow_dirty = True
        return params

    def best_action(self, params: ParamTree, state, key):
        """isdqn.py:127-135: a uniformly drawn online head, then its greedy action.  Returns a 0-d int32 CUDA tensor
        (`.item()` synchronises, like the reference's `.item()` in collect_single_sample).  The head draw is
        `jax.random.randint(key, (), 0, K)` restated on the host (isdqn_threefry_randint): the same JAX key picks the
        same head as in the reference."""
        k0, k1 = _raw_key(key)
        idx_network = int(self._lib.isdqn_threefry_randint(k0, k1, 0, self.n_bellman_iterations))
        return self.best_action_of_head(params, state, idx_network)

    def _greedy_actions_fast(self, params: ParamTree, obs: np.ndarray):
        """Greedy action of every head for ONE host observation, as host int32[1 + K]: pinned staging -> H2D -> the
        learner's own forward on a batch of one (s' aliased to s; fp32 or tensor-core path like the learner) + argmax of
        every head, replayed as one CUDA graph -> D2H; one synchronisation (the reference's `.item()`)."""
        t = self._torch
        lib = self._lib
        net = self.network
        ctx = self._context(1)
        a = ctx.get("act")
        nq = 1 + self.n_bellman_iterations
        if a is None:
            nbytes = ctx["state"].numel() * ctx["state"].element_size()
            h_obs = _lib.pinned_block(nbytes)
            h_arg = _lib.pinned_block(4 * nq).view(t.int32)
            a = ctx["act"] = {
                "h_obs": h_obs, "h_obs_np": h_obs.numpy().view(np.uint8 if ctx["state"].dtype == t.uint8 else np.float32),
                "nbytes": nbytes, "q": t.empty((2, net.final_feature), dtype=t.float32, device="cuda"),
                "d_arg": t.empty(nq, dtype=t.int32, device="cuda"), "h_arg": h_arg, "h_arg_np": h_arg.numpy(),
                "ev": _new_event(lib), "graph": None, "key": None, "warm": 0,
                "batch": _lib.Batch(ctx["state"].data_ptr(), ctx["state"].data_ptr(), ctx["action"].data_ptr(),
                                    ctx["reward"].data_ptr(), ctx["terminal"].data_ptr()),
            }
        a["h_obs_np"][...] = obs.reshape(-1)
        cur = t.cuda.current_stream()
        if a.get("fused") is None:
            # the single-kernel forward (csrc/acting.cu): cnn on uint8 frames; 0 bytes = not covered, keep the layer chain
            nb = int(lib.isdqn_act_workspace_bytes(net._net)) if net.architecture_type == "cnn" else 0
            if os.environ.get("ISDQN_ACT_FUSED", "1") == "0":
                nb = 0
            a["fused"] = t.zeros(nb, dtype=t.uint8, device="cuda") if nb > 0 else False
            if nb > 0:
                t.cuda.current_stream().synchronize()  # the zeroed barrier words precede the first launch on any stream
        if a["fused"] is not False and os.environ.get("ISDQN_ACT_MAPPED", "1") != "0":
            # no copies, no event: the kernel reads the pinned observation and writes actions + a flag the host spins on
            if "h_flag" not in a:
                a["h_flag"] = _lib.pinned_block(64).view(t.int32)
                a["seq"] = 0
            a["seq"] = (a["seq"] + 1) & 0x3FFFFFFF
            _lib.check(
                lib.isdqn_act_mapped(net._net, params.flat.data_ptr(), a["h_obs"].data_ptr(), a["q"].data_ptr(),
                                     a["h_arg"].data_ptr(), a["h_flag"].data_ptr(), a["seq"], a["fused"].data_ptr(),
                                     a["fused"].numel(), cur.cuda_stream, 5_000_000),
                "isdqn_act_mapped",
            )
            return a["h_arg_np"]
        if a["fused"] is not False:
            _lib.check(
                lib.isdqn_act_host(net._net, params.flat.data_ptr(), a["h_obs"].data_ptr(), ctx["state"].data_ptr(), a["nbytes"],
                                   a["q"].data_ptr(), a["d_arg"].data_ptr(), a["h_arg"].data_ptr(), a["fused"].data_ptr(),
                                   a["fused"].numel(), cur.cuda_stream, a["ev"]),
                "isdqn_act_host",
            )
            return a["h_arg_np"]
        side = None
        if self._use_graph and cur.cuda_stream == 0:
            if self._side_stream is None:
                self._side_stream = t.cuda.Stream()
            side = self._side_stream
            side.wait_stream(cur)
        run = side if side is not None else cur
        sp = run.cuda_stream
        _lib.check(lib.isdqn_write_async(ctx["state"].data_ptr(), a["h_obs"].data_ptr(), a["nbytes"], sp), "isdqn_write_async")
        if ctx["ws_tc"] is not None and (params.shadow is None or params.shadow_dirty):
            self._refresh_shadow(params, sp)

        def forward():
            tr = self._train_struct(ctx, params, None, 1, refresh_shadow=False)
            rc = lib.isdqn_loss_on_batch(net._net, tr, a["batch"], a["q"].data_ptr(), sp)
            if rc == 0:
                rc = lib.isdqn_argmax_heads(a["q"].data_ptr(), nq, self.n_actions, a["d_arg"].data_ptr(), sp)
            return rc

        key = (params.flat.data_ptr(), sp)
        if self._use_graph and a["graph"] is not None and a["key"] == key:
            _lib.check(lib.isdqn_graph_launch(a["graph"], sp), "isdqn_graph_launch")
        elif self._use_graph and a["warm"] >= 1:
            if a["graph"] is not None:
                lib.isdqn_graph_destroy(a["graph"])
                a["graph"] = None
            _lib.check(lib.isdqn_graph_begin(sp), "isdqn_graph_begin")
            rc = forward()
            exec_ = _lib.C.c_void_p()
            rc2 = lib.isdqn_graph_end(sp, exec_)
            _lib.check(rc, "best_action forward (capture)")
            _lib.check(rc2, "isdqn_graph_end")
            a["graph"], a["key"] = exec_, key
            _lib.check(lib.isdqn_graph_launch(a["graph"], sp), "isdqn_graph_launch")
        else:
            _lib.check(forward(), "best_action forward")
            a["warm"] += 1
        _lib.check(lib.isdqn_read_async(a["h_arg"].data_ptr(), a["d_arg"].data_ptr(), 4 * nq, sp, a["ev"]), "isdqn_read_async")
        if side is not None:
            cur.wait_stream(side)
        _lib.check(lib.isdqn_event_synchronize(a["ev"]), "isdqn_event_synchronize")
        return a["h_arg_np"]

    def best_actions(self, params: ParamTree, states, keys):
        """Batched acting for N environments at once (new functionality, the reference acts on one
        environment): states (N, *observation_dim) host array of the stored dtype, keys a sequence of N seeds.  Row i
        draws its online head from keys[i] exactly like best_action(params, states[i], keys[i]) and takes that head's
        greedy action.  One forward on N rows + the argmax of every head, one synchronisation.  Returns int32[N]."""
        t = self._torch
        lib = self._lib
        net = self.network
        states = np.ascontiguousarray(states)
        N = int(states.shape[0])
        heads = np.array([int(lib.isdqn_threefry_randint(*_raw_key(k), 0, self.n_bellman_iterations)) for k in keys],
                         dtype=np.int64)
        if heads.shape[0] != N:
            raise ValueError("best_actions needs one key per state")
        ctx = self._context(N)
        nq = 1 + self.n_bellman_iterations
        cur = t.cuda.current_stream()
        ctx["state"].copy_(t.from_numpy(states).reshape(ctx["state"].shape), non_blocking=False)
        if ctx["ws_tc"] is not None and (params.shadow is None or params.shadow_dirty):
            self._refresh_shadow(params, cur.cuda_stream)
        q = t.empty((2 * N, net.final_feature), dtype=t.float32, device="cuda")
        d_arg = t.empty(N * nq, dtype=t.int32, device="cuda")
        batch = _lib.Batch(ctx["state"].data_ptr(), ctx["state"].data_ptr(), ctx["action"].data_ptr(), ctx["reward"].data_ptr(),
                           ctx["terminal"].data_ptr())
        tr = self._train_struct(ctx, params, None, N, refresh_shadow=False)
        _lib.check(lib.isdqn_loss_on_batch(net._net, tr, batch, q.data_ptr(), cur.cuda_stream), "isdqn_loss_on_batch")
        _lib.check(lib.isdqn_argmax_heads(q.data_ptr(), N * nq, self.n_actions, d_arg.data_ptr(), cur.cuda_stream), "isdqn_argmax_heads")
        greedy = d_arg.cpu().numpy().reshape(N, nq)
        return greedy[np.arange(N), 1 + heads].astype(np.int32)

    def best_action_of_head(self, params: ParamTree, state, idx_network: int):
        t = self._torch
        net = self.network
        if not isinstance(state, t.Tensor):
            obs = np.asarray(state)
            want = np.uint8 if net.image_input else np.float32
            if obs.dtype == want and obs.size == int(np.prod(net.observation_dim)) and self._nccl_comm is None:
                # host observation of the stored dtype: the graph-replayed path; a host int32 (`.item()` works on it)
                return np.int32(self._greedy_actions_fast(params, obs)[1 + int(idx_network)])
        x, rows, is_float = net.prepare_input(state)
        assert rows == 1
        out = t.empty((), dtype=t.int32, device="cuda")
        ws = net._workspace(1)
        _lib.check(
            self._lib.isdqn_best_action(
                net._net, params.flat.data_ptr(), x.data_ptr(), is_float, int(idx_network), out.data_ptr(), ws.data_ptr(),
                ws.numel(), _lib.stream_ptr(),
            ),
            "isdqn_best_action",
        )
        return out

    def load_model(self, model_or_path) -> None:
        """Inverse of get_model: a `{"params": ...}` dict or the path of a pickle the reference's
        save_data wrote (experiments/base/utils.py:134-135; jax-array leaves are read without jax, checkpoint.py)."""
        from ..checkpoint import load_model

        load_model(self, model_or_path)

    def get_model(self):
        """isdqn.py:137-138: `{"params": params}` with host numpy leaves (pickle-able, flax-shaped)."""
        def host(node):
            return {k: host(v) if isinstance(v, dict) else v.detach().cpu().numpy() for k, v in node.items()}

        return {
            "params": {
                "params": host(self.params["params"])
            }
        }
