"""Learner timings of bench.py (policy forward @4104, PPO minibatch step @1024), three repeats; RLCA_BWD_STREAMS=0 turns
the side streams of the backward off."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device('cuda:0')
for rep in range(3):
    r = bench.time_learner(dev, 4104)
    print(json.dumps({'bwd_streams': os.environ.get('RLCA_BWD_STREAMS', '1'), 'forward_us': r['policy_forward_us']['us'],
                      'minibatch_step_us': r['ppo_minibatch_step_us']['us']}))
