"""Helpers shared by the GPU learner tests: oracle <-> product parameter exchange and the error metric."""
import numpy as np
import torch

from oracle import learner_oracle as L


def rel_err(got, want) -> float:
    """max |got - want| / max |want|  (the metric of the 1e-5 / 2e-2 parity bars, written out once)."""
    g = np.asarray(got.detach().cpu() if isinstance(got, torch.Tensor) else got, dtype=np.float64)
    w = np.asarray(want.detach().cpu() if isinstance(want, torch.Tensor) else want, dtype=np.float64)
    assert g.shape == w.shape, (g.shape, w.shape)
    denom = max(np.abs(w).max(), 1e-30)
    return float(np.abs(g - w).max() / denom)


def make_agent(seed, obs_dim, A, K, features, layer_norm, arch, lr=6.25e-5, gamma=0.99, horizon=1, eps=1.5e-4, **kw):
    from isdqn_b200.networks.isdqn import iSDQN

    return iSDQN(seed, obs_dim, A, K, features, layer_norm, False, arch, lr, gamma, horizon, 1, 10**9, adam_eps=eps, **kw)


def oracle_params_for(agent, seed, randomize=True, dtype=torch.float64):
    net = agent.network
    p = L.init_params(seed, net.architecture_type, net.observation_dim, net.features, net.final_feature, net.layer_norm, dtype)
    if randomize:
        L.randomize_small_leaves(p, seed + 1)
    # the product computes in fp32: start both sides from the SAME fp32-representable values
    for mod in p.values():
        for leaf in mod:
            mod[leaf] = mod[leaf].to(torch.float32).to(dtype)
    return p


def push_params(agent, oracle_params, tree=None):
    tree = tree if tree is not None else agent.params
    from isdqn_b200.networks.architectures.dqn import module_at

    for mod, leaves in oracle_params.items():
        for leaf, v in leaves.items():
            module_at(tree["params"], mod)[leaf] = v.detach().to(torch.float32).numpy()


def tree_to_numpy(tree):
    """{module path: {leaf: float64 ndarray}} (impala's nested modules come back under "Stack_0/Conv_1")"""
    out = {}
    for mod, leaf, v in tree.leaves():
        out.setdefault(mod, {})[leaf] = v.detach().cpu().numpy().astype(np.float64)
    return out


def batch_as_element(batch):
    from isdqn_b200.sample_collection.replay_buffer import ReplayElement

    s, a, r, s2, d = batch
    return ReplayElement(s.numpy(), a.numpy(), r.numpy(), s2.numpy(), d.numpy())
