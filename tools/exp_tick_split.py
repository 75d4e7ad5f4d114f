"""Headline tick split: graph-replayed ticks vs graph-replayed lidar-only launches (rlca_raycast at the same poses)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl_collision_avoidance_b200.stage_world import StageWorld


def graph_time(fn, n=100, reps=10):
    for i in range(10):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


if __name__ == '__main__':
    only = sys.argv[1:]                                   # optional: scenario names to run
    for scen, worlds, ar in (('stage1', 171, 1), ('stage2', 94, 2), ('circle', 41, 1)):
        if only and scen not in only:
            continue
        env = StageWorld(512, scenario=scen, num_worlds=worlds, seed=0, auto_reset=ar)
        env.reset_pose()
        acts = [torch.rand(env.N, 2, device='cuda') for _ in range(16)]
        slots = max(2, int(300e6 / (env.N * 2048)) + 1)
        ring = torch.empty(slots, env.N, 512, device='cuda')
        t_tick = graph_time(lambda i: env.control_vel(acts[i % 16], obs_out=ring[i % slots]))
        pose = env.state['pose'].clone()
        t_lidar = graph_time(lambda i: env.raycast(pose, normalise=True, out=ring[i % slots]))
        print(json.dumps({'scenario': scen, 'agents': env.N, 'tick_us': t_tick, 'lidar_only_us': t_lidar,
                          'physics_and_gap_us': t_tick - t_lidar}))
        env.close()
        del ring
