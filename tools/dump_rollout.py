"""Render a short rollout of one world to an animated GIF + trajectory .npz (offline replacement for the Stage GUI).
    python tools/dump_rollout.py --scenario stage1 --ticks 100 --out gpurun_out/stage1"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl_collision_avoidance_b200.model.net import CNNPolicy
from rl_collision_avoidance_b200.model.ppo import generate_action
from rl_collision_avoidance_b200.render import record
from rl_collision_avoidance_b200.stage_world import StageWorld


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scenario', default='stage1')
    ap.add_argument('--ticks', type=int, default=100)
    ap.add_argument('--out', default='rollout')
    ap.add_argument('--policy', default=None)
    args = ap.parse_args()
    env = StageWorld(512, scenario=args.scenario, num_worlds=1, seed=0, auto_reset={'stage1': 1, 'stage2': 2, 'circle': 1}[args.scenario])
    env.reset_pose()
    pol = CNNPolicy(max_batch=env.N)
    if args.policy:
        pol.load_state_dict(torch.load(args.policy, map_location='cuda'))
    st = [env.obs[:, None, :].repeat(1, 3, 1).contiguous(), torch.empty(env.N, 3, 512, device='cuda')]
    k = [0]

    def step(e):
        _, _, _, scaled = generate_action(e, (st[k[0] % 2], e.gs.clone()), pol, [[0, -1], [1, 1]])
        e.control_vel(scaled, stack_in=st[k[0] % 2], stack_out=st[(k[0] + 1) % 2])
        k[0] += 1
    n = record(env, step, args.ticks, gif_path=args.out + '.gif', npz_path=args.out + '.npz', every=2)
    print('wrote', args.out + '.gif', n, 'frames;', args.out + '.npz')


if __name__ == '__main__':
    main()
