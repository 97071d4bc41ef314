import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element, make_agent, oracle_params_for, push_params, rel_err, tree_to_numpy
def l2(g, w):
    g = np.asarray(g, np.float64); w = np.asarray(w.detach().cpu() if isinstance(w, torch.Tensor) else w, np.float64)
    return float(np.linalg.norm(g - w) / max(np.linalg.norm(w), 1e-30))
cfg = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch="impala")
B = int(os.environ.get("B", "32"))
agent = make_agent(92, **cfg, compute_dtype="bfloat16")
p = oracle_params_for(agent, 92); push_params(agent, p)
batch = L.make_batch(9200, B, cfg["obs_dim"], 9, "impala"); el = batch_as_element(batch)
loss, (losses, _) = agent.loss_on_batch(agent.params, el)
grads, _ = agent.grad_on_batch(agent.params, el)
gn = tree_to_numpy(grads)
_, _, og, oq, _ = L.learn_on_batch(L.clone_params(p), L.zeros_like_params(p), L.zeros_like_params(p), 0, batch, "impala", True, 9, 9, agent.gamma, 1, 0.0, 1.0)
_, _, eg, eq, _ = L.learn_on_batch(L.clone_params(p), L.zeros_like_params(p), L.zeros_like_params(p), 0, batch, "impala", True, 9, 9, agent.gamma, 1, 0.0, 1.0, emulate_bf16=True)
print("q vs f64", rel_err(agent.last_all_q_values, oq), "q vs emu", rel_err(agent.last_all_q_values, eq), "emu vs f64", rel_err(eq, oq))
for m in og:
    for k in og[m]:
        print(f"{m}.{k:7s} vs f64 {l2(gn[m][k], og[m][k]):.3e}  vs emu {l2(gn[m][k], eg[m][k]):.3e}  emu vs f64 {l2(eg[m][k].numpy(), og[m][k]):.3e}")
