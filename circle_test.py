#!/usr/bin/env python
"""Circle-swap evaluation (drop-in for /root/reference/circle_test.py:36-131): 50 robots per world on a 25 m
circle drive to their antipodes with the deterministic (mean) action of policy/stage2.pth; a robot that has
terminated keeps its angular command but gets zero linear velocity (circle_test.py:64-65).

The reference loops until ROS shuts down and records nothing; here the run stops after --steps ticks (or when
every robot has terminated) and prints the metrics the reference never computed: success rate, collision rate,
time-outs, mean steps to goal.
    python circle_test.py --num-worlds 4 --steps 600
"""
import argparse
import os

import torch

from rl_collision_avoidance_b200.circle_world import StageWorld
from rl_collision_avoidance_b200.model.net import CNNPolicy
from rl_collision_avoidance_b200.model.ppo import generate_action_no_sampling

MAX_EPISODES = 5000
LASER_BEAM = 512
LASER_HIST = 3
HORIZON = 200
GAMMA = 0.99
LAMDA = 0.95
BATCH_SIZE = 512
EPOCH = 3
COEFF_ENTROPY = 5e-4
CLIP_VALUE = 0.1
NUM_ENV = 50
OBS_SIZE = 512
ACT_SIZE = 2
LEARNING_RATE = 5e-5


def enjoy(env, policy, action_bound, max_steps):
    """The loop of /root/reference/circle_test.py:36-84, batched.  `terminal` is the terminate flag of the PREVIOUS tick
    (not a latch), exactly as the reference carries it: a robot standing on its goal stays terminal and keeps v = 0,
    a crashed robot that turns itself free moves again.  Returns (ever_terminal, first result code, tick of the first
    termination, ticks run) for the metrics."""
    env.reset_world()                                           # circle_test.py:39-40
    env.reset_pose()
    env.generate_goal_point()
    N, dev = env.N, env.device
    obs = env.get_laser_observation()
    stacks = [obs[:, None, :].repeat(1, 3, 1).contiguous(), torch.empty(N, 3, LASER_BEAM, device=dev)]
    terminal = torch.zeros(N, dtype=torch.bool, device=dev)
    ever = torch.zeros(N, dtype=torch.bool, device=dev)
    result = torch.zeros(N, dtype=torch.uint8, device=dev)
    steps_to_end = torch.zeros(N, dtype=torch.int32, device=dev)
    for step in range(1, max_steps + 1):
        k = (step - 1) % 2
        mean, scaled_action = generate_action_no_sampling(env=env, state_list=(stacks[k], env.get_local_goal(),
                                                                                env.get_self_speed()),
                                                          policy=policy, action_bound=action_bound)
        real_action = scaled_action.clone()
        real_action[terminal, 0] = 0                            # circle_test.py:64-65
        env.control_vel(real_action, stack_in=stacks[k], stack_out=stacks[1 - k])
        r, terminal, res = env.get_reward_and_terminate(step)   # :70 - the flag of THIS tick drives the next one
        terminal = terminal.clone()
        newly = terminal & ~ever
        result[newly] = res[newly]
        steps_to_end[newly] = step
        ever |= terminal
        if bool(ever.all()):
            break
    return ever, result, steps_to_end, step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--num-worlds', type=int, default=1)
    ap.add_argument('--steps', type=int, default=600)
    ap.add_argument('--policy', default='policy/stage2.pth')
    ap.add_argument('--seed', type=int, default=0)
    args = ap.parse_args()
    env = StageWorld(OBS_SIZE, index=0, num_env=NUM_ENV, num_worlds=args.num_worlds, seed=args.seed, auto_reset=0)
    action_bound = [[0, -1], [1, 1]]
    policy = CNNPolicy(frames=LASER_HIST, action_space=2, max_batch=env.N)
    if os.path.exists(args.policy):
        print('####################################')
        print('############Loading Model###########')
        print('####################################')
        policy.load_state_dict(torch.load(args.policy, map_location='cuda'))
    else:
        print('Error: Policy File Cannot Find (%s) - evaluating the randomly initialised policy' % args.policy)
    terminal, result, steps_to_end, steps = enjoy(env, policy, action_bound, args.steps)
    n = env.N
    reach = int((result == 1).sum())
    crash = int((result == 2).sum())
    tout = int((result == 3).sum())
    mean_steps = float(steps_to_end[result == 1].float().mean()) if reach else float('nan')
    print('robots %d  ticks %d  reach goal %.1f%%  crashed %.1f%%  time out %.1f%%  still running %.1f%%  '
          'mean steps to goal %.1f' % (n, steps, 100.0 * reach / n, 100.0 * crash / n, 100.0 * tout / n,
                                       100.0 * (n - reach - crash - tout) / n, mean_steps))


if __name__ == '__main__':
    main()
