#!/usr/bin/env python
"""GPU-side early-return timing of the fused tick: 64 ticks captured in a CUDA graph per RLCA_DEBUG level, so the
Python launch rate (12.5 us per call, tools/exp_phases.py) no longer hides the prologue.  Timing experiment only: with
an early return the simulation state is not advanced correctly.  Levels: see tools/exp_phases.py."""
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

TICKS = 64
env = StageWorld(512, scenario='stage1', num_worlds=171, seed=0, auto_reset=True)
env.reset_pose()
for k in env._st[0]:
    env._st[1 - env._cur][k].copy_(env._st[env._cur][k])
rng = np.random.default_rng(0)
acts = [torch.from_numpy(random_actions(rng, env.N)).cuda() for _ in range(TICKS)]
ring = torch.empty(TICKS, env.N, 512, device='cuda')
for i in range(20):                       # eager warm-up: module load, steady state
    env.control_vel(acts[i % TICKS], obs_out=ring[i % TICKS])
torch.cuda.synchronize()
for dbg in (None, '6', '7', '3', '4', '5', '1', '2', None):
    try:
        if dbg is None:
            os.environ.pop('RLCA_DEBUG', None)
        else:
            os.environ['RLCA_DEBUG'] = dbg
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(TICKS):
                env.control_vel(acts[i], obs_out=ring[i])
        os.environ.pop('RLCA_DEBUG', None)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({'exp': 'phases_graph', 'RLCA_DEBUG': dbg, 'us_per_tick': e0.elapsed_time(e1) / (5 * TICKS) * 1e3}),
              flush=True)
        del g
    except Exception as ex:               # keep going: one failing level must not lose the others
        print(json.dumps({'exp': 'phases_graph', 'RLCA_DEBUG': dbg, 'error': repr(ex)[:300]}), flush=True)
        os.environ.pop('RLCA_DEBUG', None)
