#!/usr/bin/env python
"""Timing experiments of round 1b on the headline workload (171 stage-1 worlds x 24 robots x 512 beams), one process:

  tick   device-resident tick (CUDA events, 500 ticks after 50 warm-up, obs into a 128-slot ring) for the 32- and
         48-register builds of the tick kernel (RLCA_WIDE) x CTAs per world
  e2e    rlca_env_step_host (pinned buffers, sync inside every call) for every host-chunk count, incl. the
         zero-copy experiment (-1)

One JSON line per measurement on stdout.  Numbers from this tool pick library defaults; the judged numbers come
from bench.py.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch

import __graft_entry__ as g

g.build(quiet=True)
from helpers import random_actions
from rl_collision_avoidance_b200.stage_world import StageWorld

WORLDS, BEAMS = 171, 512
dev = torch.device('cuda', 0)


def make_env(wide, ctas, worlds=WORLDS, scenario='stage1'):
    os.environ['RLCA_WIDE'] = str(wide)
    env = StageWorld(BEAMS, index=0, scenario=scenario, num_worlds=worlds, device=dev, seed=0,
                     auto_reset=2 if scenario == 'stage2' else True, ctas_per_world=ctas)
    env.reset_pose()
    return env


def time_ticks(env, acts, ring, steps=500, warm=50):
    for i in range(warm):
        env.control_vel(acts[i % 64], obs_out=ring[i % 128])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        env.control_vel(acts[i % 64], obs_out=ring[i % 128])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


def main():
    rng = np.random.default_rng(1000)
    N = WORLDS * 24
    acts = [torch.from_numpy(random_actions(rng, N)).to(dev) for _ in range(64)]
    ring = torch.empty(128, N, BEAMS, device=dev)
    for rep in range(2):                       # two passes: the second one shows the run-to-run spread
        for wide in (0, 1):
            for ctas in (0, 4):
                env = make_env(wide, ctas)
                us = time_ticks(env, acts, ring)
                print(json.dumps({'exp': 'tick', 'rep': rep, 'wide': wide, 'ctas_per_world': ctas, 'us_per_tick': us,
                                  'agent_steps_per_s': N / us * 1e6}), flush=True)
                env.close()
    # other batch sizes / scenario: which register build wins when more than 5 CTAs per SM could be resident?
    for rep in range(2):
        for scenario, worlds, R in (('stage1', 43, 24), ('stage1', 684, 24), ('stage1', 2731, 24), ('stage2', 94, 44)):
            n = worlds * R
            a2 = [torch.from_numpy(random_actions(rng, n)).to(dev) for _ in range(64)]
            r2 = torch.empty(128 if n < 20000 else 16, n, BEAMS, device=dev)
            for wide in (0, 1, 2):
                env = make_env(wide, 0, worlds, scenario)
                for i in range(20):
                    env.control_vel(a2[i % 64], obs_out=r2[i % r2.shape[0]])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(200):
                    env.control_vel(a2[i % 64], obs_out=r2[i % r2.shape[0]])
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 200 * 1e3
                print(json.dumps({'exp': 'tick_sizes', 'rep': rep, 'scenario': scenario, 'robots': n, 'wide': wide,
                                  'us_per_tick': us, 'agent_steps_per_s': n / us * 1e6}), flush=True)
                env.close()
            del a2, r2
    env = make_env(1, 0)
    a_host = [torch.from_numpy(random_actions(rng, N)).pin_memory() for _ in range(8)]
    for rep in range(2):
        for zc in (0, 2, 1):
            env.set_host_zero_copy(zc)
            for k in ((1, 2, 3) if zc != 1 else (1,)):
                env.set_host_chunks(k)
                for want_obs in (True, False):
                    if not want_obs and k != 1:
                        continue
                    for i in range(10):
                        env.step_host(a_host[i % 8], want_obs=want_obs)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for i in range(300):
                        env.step_host(a_host[i % 8], want_obs=want_obs)
                    torch.cuda.synchronize()
                    us = (time.perf_counter() - t0) / 300 * 1e6
                    print(json.dumps({'exp': 'e2e', 'rep': rep, 'zero_copy': zc, 'host_chunks': k, 'want_obs': want_obs,
                                      'us_per_step': us, 'agent_steps_per_s': N / us * 1e6}), flush=True)


if __name__ == '__main__':
    main()
