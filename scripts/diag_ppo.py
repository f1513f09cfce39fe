import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+"/tests")
import torch, numpy as np
from isaac_b200 import _lib, build
lib = _lib.load()
from test_ppo_parity import make_pair, rel_l2
from oracle import make_golden as mg
from oracle.ppo_oracle import mlp
import torch.nn.functional as F
dev = torch.device("cuda:0")
alg, ora, params = make_pair(dev, 64, 4, dict(mg.PPO_ALG, schedule="fixed"))
g = torch.Generator().manual_seed(0)
obs, cobs = torch.randn(200, 615, generator=g), torch.randn(200, 1050, generator=g)
ac = alg.actor_critic
ws = ac.workspace(200)
x = ac._as_operand(obs.to(dev), 615)
print("operand in place:", x.data_ptr() == obs.to(dev).data_ptr(), x.shape, x.stride())
out = ac._mlp_forward("actor", x, ws)
torch.cuda.synchronize()
h = obs
with torch.no_grad():
    for i, k in enumerate((0, 2, 4)):
        h = F.elu(F.linear(h, ora.params[f"actor.{k}.weight"], ora.params[f"actor.{k}.bias"]))
        got = ws["actor"]["h"][i][:, :h.shape[1]].cpu()
        print("layer", i, "rel_l2", rel_l2(got, h), "max abs", (got - h).abs().max().item(), "ones col", ws["actor"]["h"][i][:, h.shape[1]].unique())
    mu = F.linear(h, ora.params["actor.6.weight"], ora.params["actor.6.bias"])
print("mu rel_l2", rel_l2(out[:, :10].cpu(), mu), "pad cols", out[:, 10:].abs().max().item())
v = ac.evaluate(cobs.to(dev)).cpu()
with torch.no_grad():
    wv = mlp(ora.params, "critic", cobs)
print("v rel_l2", rel_l2(v, wv), (v-wv).abs().max().item(), wv.abs().mean().item())
