#!/usr/bin/env python
"""Early-return timing of the fused tick (RLCA_DEBUG, timing experiments only) on the headline workload, one process.

RLCA_DEBUG: 6 = return at kernel entry (launch overhead), 7 = after the TMA of the static tile, 3 = after phase A
(state loads, integrate) + TMA wait, 4 = after the owner-grid marking, 5 = after the collision test, 1 = after phase B
(reward/done, re-spawn, state writes; no lidar), 2 = after lidar phase 1 (walk lists), unset = the whole tick.
With an early return the simulation state is not advanced correctly; only the times are meaningful.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch

from helpers import random_actions
from rl_collision_avoidance_b200.stage_world import StageWorld

for dbg in (None, '6', '7', '3', '4', '5', '1', '2', None):
    if dbg is None:
        os.environ.pop('RLCA_DEBUG', None)
    else:
        os.environ['RLCA_DEBUG'] = dbg
    env = StageWorld(512, scenario='stage1', num_worlds=171, seed=0, auto_reset=True)
    os.environ.pop('RLCA_DEBUG', None)
    env.reset_pose()
    for k in env._st[0]:                   # both ping-pong buffers valid: early returns skip the state write
        env._st[1 - env._cur][k].copy_(env._st[env._cur][k])
    if dbg is not None:
        os.environ['RLCA_DEBUG'] = dbg
    rng = np.random.default_rng(0)
    acts = [torch.from_numpy(random_actions(rng, env.N)).cuda() for _ in range(64)]
    ring = torch.empty(128, env.N, 512, device='cuda')
    for i in range(100):
        env.control_vel(acts[i % 64], obs_out=ring[i % 128])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(1000):
        env.control_vel(acts[i % 64], obs_out=ring[i % 128])
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({'exp': 'phases', 'RLCA_DEBUG': dbg, 'us_per_tick': e0.elapsed_time(e1)}), flush=True)
    env.close()
