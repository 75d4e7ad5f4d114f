"""CUDA-event timings + achieved HBM GB/s of the stand-alone kernels the north-star names:
raycast roofline sweep (BASELINE config 5: 65536 robots x {180,360,512,1024} beams), GAE, Adam, fused tick at
other batch sizes.  Prints one JSON line per measurement (copied into profiles/)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rl_collision_avoidance_b200.model.net import Adam, CNNPolicy
from rl_collision_avoidance_b200.model.ppo import generate_train_data
from rl_collision_avoidance_b200.stage_world import StageWorld

PEAK = 6583.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def main():
    out = []
    # ---- raycast sweep (stand-alone rlca_raycast): algorithmic bytes 4*B + 12 per robot
    worlds = 2731
    for beams in (180, 360, 512, 1024):
        env = StageWorld(beams, scenario='stage1', num_worlds=worlds, seed=0, raw_beams=beams)
        env.reset_pose()
        pose = env.state['pose'].clone()
        ring = [torch.empty(env.N, beams, device='cuda') for _ in range(max(2, int(300e6 / (env.N * beams * 4)) + 1))]
        k = [0]

        def f():
            env.raycast(pose, out=ring[k[0] % len(ring)])
            k[0] += 1
        t = timeit(f, n=30)
        by = env.N * (4 * beams + 12)
        out.append({'kernel': 'rlca_raycast', 'robots': env.N, 'beams': beams, 'us': t * 1e6, 'robots_per_s': env.N / t,
                    'rays_per_s': env.N * beams / t, 'alg_bytes': by, 'GBps': by / t / 1e9, 'frac_of_measured_hbm': by / t / 1e9 / PEAK})
        env.close()
        del env, ring
    # ---- fused tick at other sizes
    for worlds in (43, 171, 684):
        env = StageWorld(512, scenario='stage1', num_worlds=worlds, seed=0, auto_reset=1)
        env.reset_pose()
        acts = [torch.rand(env.N, 2, device='cuda') for _ in range(16)]
        ring = torch.empty(max(2, int(300e6 / (env.N * 2048)) + 1), env.N, 512, device='cuda')
        k = [0]

        def f():
            env.control_vel(acts[k[0] % 16], obs_out=ring[k[0] % ring.shape[0]])
            k[0] += 1
        t = timeit(f, n=200, warm=20)
        by = env.N * (4 * 512 + 96)
        out.append({'kernel': 'fused tick', 'robots': env.N, 'beams': 512, 'us': t * 1e6, 'agent_steps_per_s': env.N / t,
                    'alg_bytes': by, 'GBps': by / t / 1e9, 'frac_of_measured_hbm': by / t / 1e9 / PEAK})
        env.close()
        del env, ring
    # ---- fused tick on the other scenarios (stage 2: 200 x 200 map + polygons, group-synchronous; circle: global-grid path)
    for scen, worlds, ar in (('stage2', 94, 2), ('circle', 41, 1)):
        env = StageWorld(512, scenario=scen, num_worlds=worlds, seed=0, auto_reset=ar)
        env.reset_pose()
        acts = [torch.rand(env.N, 2, device='cuda') for _ in range(16)]
        k = [0]

        def g():
            env.control_vel(acts[k[0] % 16])
            k[0] += 1
        t = timeit(g, n=100, warm=10)
        by = env.N * (4 * 512 + 96)
        out.append({'kernel': 'fused tick ' + scen, 'robots': env.N, 'beams': 512, 'us': t * 1e6, 'agent_steps_per_s': env.N / t,
                    'alg_bytes': by, 'GBps': by / t / 1e9, 'frac_of_measured_hbm': by / t / 1e9 / PEAK})
        env.close()
        del env
    # ---- GAE: T x N x (4 r + 4 v + 1 d + 4 v[t+1 from L2] ... ) algorithmic = 17 bytes per element (r, v, d in; target, adv out)
    for (T, N) in ((128, 4104), (128, 16416)):
        r = torch.randn(T, N, device='cuda'); v = torch.randn(T, N, device='cuda'); lv = torch.randn(N, device='cuda')
        d = torch.rand(T, N, device='cuda') < 0.05
        t = timeit(lambda: generate_train_data(r, 0.99, v, lv, d, 0.95), n=50)
        by = T * N * 17
        out.append({'kernel': 'gae_kernel', 'T': T, 'N': N, 'us': t * 1e6, 'alg_bytes': by, 'GBps': by / t / 1e9,
                    'frac_of_measured_hbm': by / t / 1e9 / PEAK})
    # ---- Adam: 28 bytes per parameter (p, g, m, v read; p, m, v written)
    pol = CNNPolicy(max_batch=8)
    opt = Adam(pol.parameters(), lr=5e-5)
    pol.grad.normal_()
    t = timeit(lambda: opt.step(), n=100)
    by = pol.flat_size * 28
    out.append({'kernel': 'adam_kernel', 'params': pol.flat_size, 'us': t * 1e6, 'alg_bytes': by, 'GBps': by / t / 1e9,
                'frac_of_measured_hbm': by / t / 1e9 / PEAK})
    for o in out:
        print(json.dumps(o))


if __name__ == '__main__':
    main()
