"""Launch the fused rollout-time policy step a few times at the bench shape (for ncu captures)."""
import torch
import pufferlib_b200.vector as pvec
from pufferlib_b200 import models
from pufferlib_b200.environments import ocean
from pufferlib_b200.frameworks import cleanrl

dev = torch.device('cuda')
vec = pvec.make(ocean.env_creator('breakout'), num_envs=16384, backend=pvec.B200)
pol = cleanrl.Policy(models.Default(vec.driver_env).to(dev), fused_sample=True, seed=1).to(dev)
x = torch.randn(16384, 128, device=dev)
with torch.no_grad():
    for _ in range(4):
        pol(x)
torch.cuda.synchronize()
print('ok', int(pol._counter.item()))
