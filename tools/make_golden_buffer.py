"""Golden vectors for transform_buffer from the reference's OWN model/ppo.py.

Run HERE (where /root/reference exists): python tools/make_golden_buffer.py
The reference's rollout buffer is a list, per step, of ([ [obs_stack, goal, speed] per robot ], a, r, d, logprob, v)
(ppo_stage1.py:95-96); model/ppo.py:22-54 stacks it into eight arrays.  This script builds such a buffer from seeded
numpy arrays, runs the reference function unmodified and stores inputs and outputs in tests/golden/buffer_golden.npz;
tests/test_learner_host.py feeds the same steps to the library's transform_buffer (batched tensors per step)."""
import builtins
import functools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('RLCA_REFERENCE', '/root/reference')


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    builtins.reduce = functools.reduce
    cwd = os.getcwd()
    os.chdir('/tmp')                                         # model/ppo.py creates ./log/<host>/ at import
    sys.path.insert(0, REF)
    from model import ppo as ref_ppo
    os.chdir(cwd)
    rs = np.random.RandomState(77)
    T, N, B = 5, 6, 16
    obs = rs.uniform(-0.5, 0.5, (T, N, 3, B)).astype(np.float32)
    goal = rs.uniform(-5, 5, (T, N, 2)).astype(np.float32)
    speed = rs.uniform(-1, 1, (T, N, 2)).astype(np.float32)
    a = rs.uniform(-1, 1, (T, N, 2)).astype(np.float32)
    r = rs.uniform(-1, 1, (T, N)).astype(np.float32)
    d = (rs.uniform(0, 1, (T, N)) < 0.3)
    lp = rs.uniform(-2, 0, (T, N, 1)).astype(np.float32)
    v = rs.uniform(-1, 1, (T, N, 1)).astype(np.float32)
    buff = []
    for t in range(T):
        state_list = [[obs[t, i], goal[t, i], speed[t, i]] for i in range(N)]
        buff.append((state_list, a[t], r[t], d[t], lp[t], v[t]))
    names = ('s', 'goal', 'speed', 'a', 'r', 'd', 'l', 'v')
    out = {'in_obs': obs, 'in_goal': goal, 'in_speed': speed, 'in_a': a, 'in_r': r, 'in_d': d, 'in_l': lp, 'in_v': v}
    for k, arr in zip(names, ref_ppo.transform_buffer(buff)):
        out['out_' + k] = np.asarray(arr)
    path = os.path.join(ROOT, 'tests', 'golden', 'buffer_golden.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items() if k.startswith('out_')})


if __name__ == '__main__':
    main()
