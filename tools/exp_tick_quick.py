#!/usr/bin/env python
"""Quick A/B timing of the device-resident tick on the headline workload and three other sizes (CUDA events)."""
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

dev = torch.device('cuda', 0)
rng = np.random.default_rng(1000)
for rep in range(2):
    for scenario, worlds, R, beams in (('stage1', 171, 24, 512), ('stage1', 43, 24, 512), ('stage1', 684, 24, 512),
                                       ('stage2', 94, 44, 512), ('stage1', 171, 24, 180), ('stage1', 171, 24, 1024)):
        n = worlds * R
        acts = [torch.from_numpy(random_actions(rng, n)).to(dev) for _ in range(64)]
        ring = torch.empty(128 if n < 20000 else 16, n, beams, device=dev)
        env = StageWorld(beams, index=0, scenario=scenario, num_worlds=worlds, device=dev, seed=0,
                         auto_reset=2 if scenario == 'stage2' else True, raw_beams=beams if beams != 180 else 512)
        env.reset_pose()
        for i in range(50):
            env.control_vel(acts[i % 64], obs_out=ring[i % ring.shape[0]])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(500):
            env.control_vel(acts[i % 64], obs_out=ring[i % ring.shape[0]])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 500 * 1e3
        print(json.dumps({'exp': 'tick', 'rep': rep, 'scenario': scenario, 'robots': n, 'beams': beams, 'us_per_tick': us,
                          'agent_steps_per_s': n / us * 1e6}), flush=True)
        env.close()
        del acts, ring
