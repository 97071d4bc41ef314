"""Debug helper (GPU): per-layer comparison of the bf16 tensor-core forward against torch fp32 ops."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import learner_oracle as L
from tests.learner_utils import make_agent, oracle_params_for, push_params, batch_as_element

cfg = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch="cnn")
B = 32
agent = make_agent(1, **cfg, compute_dtype="bfloat16")
p = oracle_params_for(agent, 1)
push_params(agent, p)
batch = L.make_batch(100, B, cfg["obs_dim"], 9, "cnn")
el = batch_as_element(batch)
agent.loss_on_batch(agent.params, el)
torch.cuda.synchronize()
ctx = agent._ctx[B]
ws = ctx["ws_tc"]
rows = 2 * B
x = torch.cat((batch[0], batch[3])).cuda().float() / 255.0
x = x.to(torch.bfloat16).float()
off = 0
geo = L.CONV_GEOMETRY
for i, (k, s) in enumerate(geo):
    w = p[f"Conv_{i}"]["kernel"].cuda().float().to(torch.bfloat16).float()
    b = p[f"Conv_{i}"]["bias"].cuda().float()
    _, a, bb = L.same_padding(x.shape[1], k, s); _, c, d = L.same_padding(x.shape[2], k, s)
    y = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (c, d, a, bb)), w.permute(3, 2, 0, 1), stride=s).permute(0, 2, 3, 1) + b
    g = p[f"LayerNorm_{i}"]["scale"].cuda().float(); be = p[f"LayerNorm_{i}"]["bias"].cuda().float()
    y = L.layer_norm_lastdim(y, g, be)
    x = torch.relu(y).to(torch.bfloat16).float()
    n = x.numel()
    got = ws[off: off + 2 * n].view(torch.bfloat16).float().view(x.shape)
    err = (got - x).abs().max().item() / x.abs().max().item()
    bad = ((got - x).abs() > 0.05 * x.abs().max()).float().mean().item()
    print(f"layer {i}: shape {tuple(x.shape)} rel err {err:.3e} frac bad {bad:.4f}")
    if bad > 0:
        idx = ((got - x).abs() > 0.05 * x.abs().max()).nonzero()
        print("  first bad idx", idx[:5].tolist(), "rows(img) hist", torch.bincount(idx[:, 0], minlength=rows)[:8].tolist())
        print("  oy hist", torch.bincount(idx[:, 1]).tolist()[:25])
        print("  ox hist", torch.bincount(idx[:, 2]).tolist()[:25])
    print("  got ", got[0, 5, 5, :6].tolist())
    print("  want", x[0, 5, 5, :6].tolist())
    off = (off + 2 * n + 255) // 256 * 256
    x = got.clone()  # isolate the next layer: feed it what the GPU produced
