"""One forward + loss + backward + Adam at the given batch sizes (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_collision_avoidance_b200 import _lib
from rl_collision_avoidance_b200.model.net import CNNPolicy, Adam, _ptr

pol = CNNPolicy(max_batch=4104)
opt = Adam(pol.parameters(), lr=5e-5)
lib = pol.lib
for nb in [int(a) for a in sys.argv[1:]] or [4104, 1024]:
    obs = torch.rand(nb, 1536, device='cuda') - 0.5
    gs = torch.rand(nb, 4, device='cuda')
    v = torch.empty(nb, device='cuda'); mean = torch.empty(nb, 2, device='cuda')
    act = torch.rand(nb, 2, device='cuda'); lp = torch.rand(nb, device='cuda') - 1; adv = torch.randn(nb, device='cuda')
    tgt = torch.randn(nb, device='cuda'); losses = torch.zeros(3, device='cuda')
    ws, st = pol._workspace(nb), pol._stream()
    for rep in range(2):
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push(f'nb{nb}_rep{rep}')
        _lib.check(lib.rlca_policy_forward(ws, _ptr(pol.flat), _ptr(obs), _ptr(gs), nb, _ptr(v), _ptr(mean), st))
        _lib.check(lib.rlca_ppo_loss_fwd_bwd(ws, _ptr(pol.flat), _ptr(v), _ptr(mean), _ptr(act), _ptr(lp), _ptr(adv), _ptr(tgt), nb, 0.1, 5e-4, 20.0, _ptr(losses), st))
        _lib.check(lib.rlca_policy_backward(ws, _ptr(pol.flat), _ptr(obs), _ptr(gs), nb, _ptr(pol.grad), st))
        opt.step()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
