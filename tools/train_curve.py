"""Short stage-1 training run that prints the learning curve (does PPO actually learn on the new env?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from rl_collision_avoidance_b200.model.net import Adam, CNNPolicy
from rl_collision_avoidance_b200.stage_world1 import StageWorld
from rl_collision_avoidance_b200.trainer import run

updates = int(sys.argv[1]) if len(sys.argv) > 1 else 100
worlds = int(sys.argv[2]) if len(sys.argv) > 2 else 43
env = StageWorld(512, index=0, num_env=24, num_worlds=worlds, seed=0, auto_reset=1)
policy = CNNPolicy(frames=3, action_space=2, seed=0, max_batch=max(1024, env.N))
opt = Adam(policy.parameters(), lr=5e-5)
hp = dict(HORIZON=128, GAMMA=0.99, LAMDA=0.95, BATCH_SIZE=1024, EPOCH=2, COEFF_ENTROPY=5e-4, CLIP_VALUE=0.1, NUM_ENV=24,
          OBS_SIZE=512, ACT_SIZE=2, LASER_HIST=3, MAX_EPISODES=10 ** 9)
stats = run(env=env, policy=policy, policy_path=None, action_bound=[[0, -1], [1, 1]], optimizer=opt, hp=hp, stage=1,
            max_updates=updates)
for i in range(0, len(stats), max(1, len(stats) // 12)):
    ch = stats[i:i + max(1, len(stats) // 12)]
    print('updates %3d-%3d  mean ep reward %7.2f  success %5.1f%%  episodes %5d  agent-steps/s %8.0f  losses %s' % (
        ch[0]['update'], ch[-1]['update'], np.nanmean([c['mean_ep_reward'] for c in ch]),
        100 * np.nanmean([c['success_rate'] for c in ch]), sum(c['episodes'] for c in ch),
        np.mean([c['agent_steps_per_s'] for c in ch]), np.round(ch[-1]['losses'], 3)))
