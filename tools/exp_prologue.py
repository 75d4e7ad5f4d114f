import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from helpers import random_actions
from rl_collision_avoidance_b200.stage_world import StageWorld
for ar in ((1, 0) if os.environ.get('RLCA_DEBUG') in ('0', '1') else (1,)):
    env = StageWorld(512, scenario='stage1', num_worlds=171, seed=0, auto_reset=ar)
    env.reset_pose()
    rng = np.random.default_rng(0)
    acts = [torch.from_numpy(random_actions(rng, env.N)).cuda() for _ in range(64)]
    ring = torch.empty(128, env.N, 512, device='cuda')
    for i in range(100): env.control_vel(acts[i % 64], obs_out=ring[i % 128])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(1000): env.control_vel(acts[i % 64], obs_out=ring[i % 128])
    e1.record(); torch.cuda.synchronize()
    print('RLCA_DEBUG', os.environ.get('RLCA_DEBUG'), 'auto_reset', ar, round(e0.elapsed_time(e1), 2), 'us/tick',
          'resets/tick', float(env.flags[:, 3].float().sum()))
