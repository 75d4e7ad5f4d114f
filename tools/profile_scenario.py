"""A few ticks of a scenario at its BASELINE size, for ncu:  python tools/profile_scenario.py circle 41 1 [ticks]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl_collision_avoidance_b200.stage_world import StageWorld

scen, worlds, ar = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ticks = int(sys.argv[4]) if len(sys.argv) > 4 else 12
env = StageWorld(512, scenario=scen, num_worlds=worlds, seed=0, auto_reset=ar)
env.reset_pose()
acts = [torch.rand(env.N, 2, device='cuda') for _ in range(4)]
ring = [torch.empty(env.N, 512, device='cuda') for _ in range(8)]
for t in range(ticks):
    env.control_vel(acts[t % 4], obs_out=ring[t % 8])
torch.cuda.synchronize()
print('done', env.launch_count)
