"""Debug: conv-tower gradient differences, tcgen05 backward vs CUDA-core backward; run-to-run determinism."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from rl_collision_avoidance_b200 import _lib
from rl_collision_avoidance_b200.model.net import CNNPolicy, _ptr
from golden_inputs import synthetic_state_dict
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for nb in [int(a) for a in sys.argv[1:]] or [37, 333]:
    rs = np.random.RandomState(100 + nb)
    x = T(rs.rand(nb, 1536).astype(np.float32) - 0.5); gs = T(rs.randn(nb, 4).astype(np.float32))
    act = T(rs.rand(nb, 2).astype(np.float32)); lp = T(rs.uniform(-1.5, 0.5, nb).astype(np.float32))
    adv = T(rs.standard_normal(nb).astype(np.float32)); tgt = T(rs.uniform(-3, 3, nb).astype(np.float32))
    grads = {}
    for mode in (1, 2, 1):
        pol = CNNPolicy(max_batch=nb); pol.set_tensor_cores(mode)
        pol.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic_state_dict().items()})
        v, mean = pol.forward_values(x, gs)
        losses = torch.zeros(3, device='cuda')
        lib, ws, st = pol.lib, pol._workspace(nb), pol._stream()
        _lib.check(lib.rlca_ppo_loss_fwd_bwd(ws, _ptr(pol.flat), _ptr(v), _ptr(mean), _ptr(act), _ptr(lp), _ptr(adv), _ptr(tgt), nb, 0.1, 5e-4, 20.0, _ptr(losses), st))
        _lib.check(lib.rlca_policy_backward(ws, _ptr(pol.flat), _ptr(x), _ptr(gs), nb, _ptr(pol.grad), st))
        torch.cuda.synchronize()
        g = {k: t.cpu().numpy().copy() for k, t in pol.grad_views.items()}
        if mode in grads and mode == 1:
            same = all(np.array_equal(g[k], grads[1][k]) for k in g)
            print(f'nb={nb}: tensor-core backward run-to-run bit-identical: {same}')
        grads[mode] = g
        f = [pol.features(t, nb).cpu().numpy() for t in range(2)]
        grads[('f', mode)] = f
    flips = [int(((grads[('f', 1)][t] > 0) != (grads[('f', 2)][t] > 0)).sum()) for t in range(2)]
    print(f'nb={nb}: relu(conv2) mask flips between tc / fp32 forward: {flips}')
    for k in grads[1]:
        if '_fea_cv' in k:
            r, got = grads[2][k], grads[1][k]
            e = np.abs(got - r)
            print(f'  {k:22s} scale {np.abs(r).max():.3e}  max err {e.max():.3e}  rel {e.max() / np.abs(r).max():.2e}  mean err {e.mean():.2e}')
