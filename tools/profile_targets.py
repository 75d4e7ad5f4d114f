"""A few launches of the stand-alone kernels the north-star names (raycast alone, GAE, Adam, PPO loss) for ncu:
  ncu --set full --clock-control none --import-source on -k regex:'rlca_world_kernel|gae_kernel|adam_kernel|ppo_loss_kernel' \
      -o gpurun_out/r2_targets python tools/profile_targets.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl_collision_avoidance_b200 import _lib
from rl_collision_avoidance_b200.model.net import Adam, CNNPolicy, _ptr
from rl_collision_avoidance_b200.model.ppo import generate_train_data
from rl_collision_avoidance_b200.stage_world import StageWorld

which = sys.argv[1:] or ['raycast', 'gae', 'adam', 'loss']
if 'raycast' in which:
    env = StageWorld(512, scenario='stage1', num_worlds=2731, seed=0)         # BASELINE config 5: 65 544 robots x 512 beams
    env.reset_pose()
    pose = env.state['pose'].clone()
    outs = [torch.empty(env.N, 512, device='cuda') for _ in range(3)]
    for o in outs:
        env.raycast(pose, out=o)
    torch.cuda.synchronize()
    env.close()
if 'gae' in which:
    T, N = 128, 4104
    r, v, lv = torch.randn(T, N, device='cuda'), torch.randn(T, N, device='cuda'), torch.randn(N, device='cuda')
    d = torch.rand(T, N, device='cuda') < 0.05
    for _ in range(3):
        generate_train_data(r, 0.99, v, lv, d, 0.95)
    torch.cuda.synchronize()
if 'adam' in which or 'loss' in which:
    pol = CNNPolicy(max_batch=1024)
    opt = Adam(pol.parameters(), lr=5e-5)
    pol.grad.normal_()
    if 'adam' in which:
        for _ in range(3):
            opt.step()
    if 'loss' in which:
        nb = 1024
        v, mean = torch.randn(nb, device='cuda'), torch.rand(nb, 2, device='cuda')
        act, lp = torch.rand(nb, 2, device='cuda'), torch.rand(nb, device='cuda') - 1
        adv, tgt, losses = torch.randn(nb, device='cuda'), torch.randn(nb, device='cuda'), torch.zeros(3, device='cuda')
        ws, st = pol._workspace(nb), pol._stream()
        for _ in range(3):
            _lib.check(pol.lib.rlca_ppo_loss_fwd_bwd(ws, _ptr(pol.flat), _ptr(v), _ptr(mean), _ptr(act), _ptr(lp), _ptr(adv),
                                                     _ptr(tgt), nb, 0.1, 5e-4, 20.0, _ptr(losses), st))
    torch.cuda.synchronize()
print('done')
