"""Time the learner kernels (CUDA events): policy forward at rollout size, one PPO minibatch step at 1024."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_collision_avoidance_b200 import _lib
from rl_collision_avoidance_b200.model.net import CNNPolicy, Adam, _ptr


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    pol = CNNPolicy(max_batch=4104)
    opt = Adam(pol.parameters(), lr=5e-5)
    lib = pol.lib
    for nb in (4104, 1024):
        obs = torch.rand(nb, 1536, device='cuda') - 0.5
        gs = torch.rand(nb, 4, device='cuda')
        v = torch.empty(nb, device='cuda'); mean = torch.empty(nb, 2, device='cuda')
        act = torch.rand(nb, 2, device='cuda'); lp = torch.rand(nb, device='cuda') - 1; adv = torch.randn(nb, device='cuda')
        tgt = torch.randn(nb, device='cuda'); losses = torch.zeros(3, device='cuda')
        ws, st = pol._workspace(nb), pol._stream()
        fwd = lambda: _lib.check(lib.rlca_policy_forward(ws, _ptr(pol.flat), _ptr(obs), _ptr(gs), nb, _ptr(v), _ptr(mean), st))
        loss = lambda: _lib.check(lib.rlca_ppo_loss_fwd_bwd(ws, _ptr(pol.flat), _ptr(v), _ptr(mean), _ptr(act), _ptr(lp), _ptr(adv), _ptr(tgt), nb, 0.1, 5e-4, 20.0, _ptr(losses), st))
        bwd = lambda: _lib.check(lib.rlca_policy_backward(ws, _ptr(pol.flat), _ptr(obs), _ptr(gs), nb, _ptr(pol.grad), st))
        t_f = timeit(fwd)
        t_l = timeit(loss)
        t_b = timeit(bwd)
        t_a = timeit(lambda: opt.step())
        fl = 6.39e6 * nb
        print(f'nb={nb}: forward {t_f:.1f} us ({fl/t_f/1e6:.2f} TFLOP/s)  loss {t_l:.1f} us  backward {t_b:.1f} us '
              f'({2*fl/t_b/1e6:.2f} TFLOP/s)  adam {t_a:.1f} us')


if __name__ == '__main__':
    main()
