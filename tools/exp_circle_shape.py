"""Circle tick vs viewers per lidar CTA (ctas_per_world hint 50 / 25 / 17 / 13 -> 1 / 2 / 3 / 4 viewers)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl_collision_avoidance_b200.stage_world import StageWorld
from tools.exp_tick_split import graph_time

for s in (0, 50, 25, 17, 13):
    env = StageWorld(512, scenario='circle', num_worlds=41, seed=0, auto_reset=1, ctas_per_world=s)
    env.reset_pose()
    acts = [torch.rand(env.N, 2, device='cuda') for _ in range(16)]
    ring = torch.empty(64, env.N, 512, device='cuda')
    t = graph_time(lambda i: env.control_vel(acts[i % 16], obs_out=ring[i % 64]), n=60, reps=5)
    print(json.dumps({'ctas_per_world_hint': s, 'tick_us': t}))
    env.close()
