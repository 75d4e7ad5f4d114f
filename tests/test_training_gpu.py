"""End-to-end smoke of the trainers on the GPU: a couple of PPO updates on tiny batches must run, log in the
reference's formats and move the parameters."""
import logging

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class ListHandler(logging.Handler):
    def __init__(self):
        super().__init__()
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


@pytest.mark.parametrize('stage', [1, 2])
def test_trainer_runs_and_logs(built, stage, tmp_path):
    from rl_collision_avoidance_b200.model.net import Adam, CNNPolicy
    from rl_collision_avoidance_b200.trainer import run
    if stage == 1:
        from rl_collision_avoidance_b200.stage_world1 import StageWorld
        num_env, bs, ep = 24, 256, 2
    else:
        from rl_collision_avoidance_b200.stage_world2 import StageWorld
        num_env, bs, ep = 44, 128, 2
    env = StageWorld(512, index=0, num_env=num_env, num_worlds=2, seed=1, auto_reset=stage)
    policy = CNNPolicy(frames=3, action_space=2, seed=1, max_batch=max(bs, env.N))
    before = policy.flat.clone()
    opt = Adam(policy.parameters(), lr=5e-5)
    hp = dict(HORIZON=32, GAMMA=0.99, LAMDA=0.95, BATCH_SIZE=bs, EPOCH=ep, COEFF_ENTROPY=5e-4, CLIP_VALUE=0.1,
              NUM_ENV=num_env, OBS_SIZE=512, ACT_SIZE=2, LASER_HIST=3, MAX_EPISODES=5000)
    lg = logging.getLogger(f'test_out_{stage}')
    lg.setLevel(logging.INFO)
    h = ListHandler()
    lg.addHandler(h)
    stats = run(env=env, policy=policy, policy_path=str(tmp_path), action_bound=[[0, -1], [1, 1]], optimizer=opt, hp=hp,
                logger=lg, logger_cal=None, stage=stage, max_updates=3, save_every=2)
    assert len(stats) == 3 and all(np.isfinite(s['losses']).all() for s in stats)
    assert float((policy.flat - before).abs().max()) > 0
    assert torch.isfinite(policy.flat).all()
    assert any(l.startswith('Env ') and 'Goal (' in l for l in h.lines), h.lines[:3]
    import os
    saved = 'Stage1_2' if stage == 1 else 'stage2_2.pth'
    assert os.path.exists(tmp_path / saved)
    sd = torch.load(tmp_path / saved)
    assert len(sd) == 23 and sd['act_fc1.weight'].shape == (256, 4096)


def test_optimizer_state_roundtrip(built, tmp_path):
    """What the reference never saved (ppo_stage1.py:122-126): Adam moments + step survive a save/load."""
    from rl_collision_avoidance_b200.model.net import Adam, CNNPolicy
    pol = CNNPolicy(seed=3, max_batch=8)
    opt = Adam(pol.parameters(), lr=5e-5)
    gen = torch.Generator(device='cuda').manual_seed(1)
    for _ in range(3):
        pol.grad.copy_(torch.randn(pol.flat_size, device='cuda', generator=gen) * 0.01)
        opt.step()
    torch.save(pol.state_dict(), tmp_path / 'w')
    torch.save({'optimizer': opt.state_dict()}, tmp_path / 'w.trainer')
    pol2 = CNNPolicy(seed=9, max_batch=8)
    opt2 = Adam(pol2.parameters(), lr=5e-5)
    pol2.load_state_dict(torch.load(tmp_path / 'w'))
    opt2.load_state_dict(torch.load(tmp_path / 'w.trainer')['optimizer'])
    g = torch.randn(pol.flat_size, device='cuda', generator=gen) * 0.01
    for p_, o_ in ((pol, opt), (pol2, opt2)):
        p_.grad.copy_(g)
        o_.step()
    assert opt2.step_count == 4
    # compare the named tensors: the 32-float alignment padding between them is not part of a state_dict
    for (k, a), (_, b) in zip(pol.named_parameters(), pol2.named_parameters()):
        assert torch.equal(a, b), k
