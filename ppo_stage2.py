#!/usr/bin/env python
"""Stage-2 trainer (drop-in for /root/reference/ppo_stage2.py): testenv map + polygon obstacles, 44 robots per
world, group-synchronous episodes, filtered idle transitions, BATCH_SIZE 512, EPOCH 4, drop_last.
Curriculum: copy a stage-1 checkpoint to policy/stage2.pth (README.md:18 of the reference)."""
from ppo_stage1 import main
from rl_collision_avoidance_b200.stage_world2 import StageWorld

NUM_ENV = 44
BATCH_SIZE = 512
EPOCH = 4

if __name__ == '__main__':
    main(stage=2, world_cls=StageWorld, num_env=NUM_ENV, batch_size=BATCH_SIZE, epoch=EPOCH, ckpt='stage2.pth')
