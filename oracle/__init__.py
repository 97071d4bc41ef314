"""oracle/ — TEST INFRASTRUCTURE, NOT PRODUCT.

CPU restatements of the reference's (theovincent/iS-DQN, package `slimdqn`) learner hot path, used only as
the checker by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs.  Nothing under `is-dqn_b200/` imports this package: the product path is the CUDA library and fails
loudly when it is missing.

Pinning status:
  * sum tree, samplers, replay buffer ........ PINNED: bit-exact against the unmodified reference modules
    (imported under `oracle/refshim.py`, 19 reference tests pass) and against the committed fixtures in
    `tests/golden/replay_*.npz` produced by `oracle/make_golden.py`.
  * NumPy Generator draws (PCG64, Lemire) .... PINNED against numpy itself (numpy ships on the GPU box).
  * learner numerics (conv/LN/Dense/Adam) .... PARITY UNPINNED: jax/flax/optax are not installable here,
    the reference's own tests for this part hold no golden numbers (tests/test_isdqn.py is self-consistency
    with a random seed).  The restatement follows flax 0.10.2 / optax 0.2.4 published semantics
     and re-asserts the reference's four algebraic tests.
"""
