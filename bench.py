#!/usr/bin/env python
"""bench.py — agent-steps/s of the fused simulator tick at 4096 robots x 512 beams.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (under torchrun for N>1)
prints ONE JSON line on rank 0.  A "step" is one fused tick (integrate + collide + 512-beam
raycast + reward/done + obs) over the per-GPU agent batch: 171 stage-1 worlds x 24 robots =
4104 agents.  Scaling is weak: every GPU gets its own 171 worlds, no
data-path collective (worlds are independent).

  value        whole-job agent-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the host-buffer C-ABI call (actions read from and obs/reward/flags/gs
               written to pinned host memory over PCIe by the tick kernel itself, sync) — the reference-facing call
  roofline     algorithmic bytes (4*B+96 per agent-step) / measured launch time
               vs the measured HBM copy peak (MEASURED_PEAKS.json, else the 6650 GB/s fallback)
  cpu_baseline the CPU oracle (port of the reference semantics) on the host cores, bounded sample

`--impl reference` times the reference's CPU path.  The literal Stage+ROS+mpi4py stack
cannot run here, so this is the oracle port with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

WORLDS_PER_GPU = 171
ROBOTS = 24
BEAMS = 512
METRIC = 'agent-steps/s @4096 robots x 512 beams (fused env tick)'


def alg_bytes(beams):           # the figure builder and judge share
    return 4 * beams + 96


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            parts = [x.strip() for x in r.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def _cpu_quota():
    """CPUs the cgroup lets this process use (cpu.max), or None."""
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            q, per = f.read().split()
        if q != 'max':
            return max(1, int(float(q) / float(per) + 0.5))
    except Exception:
        pass
    return None


def run_cpu(steps, warmup, budget_s, threads=None):
    """Time the oracle port on the host cores.  Returns (agent_steps_per_s, cores, sample_desc, ms_per_tick).

    The fair CPU arm is the FASTEST configuration of the host: torchrun exports OMP_NUM_THREADS=1, and under a cgroup CPU
    quota "all logical CPUs" can be ~10x slower than fewer threads - but only after the quota's burst allowance is used
    up, so a probe of a few ticks picks the wrong count.  Every candidate thread count is therefore run for a sustained
    slice (>= 1.5 s and >= 3 ticks of the full workload), the best sustained rate picks the count, and the timed sample is
    repeated with the runner-up if it falls far below the probe (throttling kicked in)."""
    import numpy as np
    from oracle import oracle as orc_mod
    from helpers import make_pair, random_actions
    _, _, orc = make_pair('stage1', num_worlds=WORLDS_PER_GPU, beams=BEAMS, auto_reset=True, seed=0, gpu=False)
    orc.reset_world()
    orc.reset_pose()
    rng = np.random.default_rng(0)
    acts = [random_actions(rng, orc.N) for _ in range(8)]
    ncpu = len(os.sched_getaffinity(0))
    quota = _cpu_quota()
    if threads:
        cands = [threads]
    else:
        cands = {max(1, ncpu), max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 32), min(ncpu, 16)}
        if quota:
            cands |= {min(ncpu, quota), min(ncpu, max(1, quota // 2))}
        cands = sorted(cands, reverse=True)
    slice_s = max(1.0, min(2.5, 0.3 * budget_s / len(cands)))
    probe = []
    for c in cands:
        orc_mod.set_threads(c)
        orc.step(acts[0])
        n, t0 = 0, time.perf_counter()
        while n < 3 or time.perf_counter() - t0 < slice_s:
            orc.step(acts[n % 8])
            n += 1
        probe.append(((time.perf_counter() - t0) / n, c))
    probe.sort()
    best = None
    for t_tick, cores in probe[:2]:
        orc_mod.set_threads(cores)
        worlds = WORLDS_PER_GPU
        total = (steps + warmup) * t_tick
        share = 0.35 * budget_s
        o = orc
        if total > share:
            worlds = min(WORLDS_PER_GPU, max(cores, int(WORLDS_PER_GPU * share / total)))
            _, _, o = make_pair('stage1', num_worlds=worlds, beams=BEAMS, auto_reset=True, seed=0, gpu=False)
            o.reset_world()
            o.reset_pose()
        a = [x[:o.N] for x in acts]
        for i in range(warmup):
            o.step(a[i % len(a)])
        t0 = time.perf_counter()
        for i in range(steps):
            o.step(a[i % len(a)])
        dt = time.perf_counter() - t0
        val = o.N * steps / dt
        sample = f'{worlds} of {WORLDS_PER_GPU} stage-1 worlds x {ROBOTS} robots x {BEAMS} beams, {steps} ticks, ' \
                 f'OpenMP over worlds, {cores} threads (best sustained of {sorted(c for _, c in probe)} threads' \
                 f'{", cgroup quota %d CPUs" % quota if quota else ""})'
        if best is None or val > best[0]:
            best = (val, cores, sample, dt / steps * 1e3)
        if val > 0.6 * WORLDS_PER_GPU * ROBOTS / probe[0][0]:      # the sample reproduced the probe: done
            break
    return best


def time_learner(dev, n_agents):
    """Side measurement (outside the timed env region, CUDA events): the policy forward at the rollout batch and one
    PPO minibatch step (forward, loss, backward, Adam) at the reference's batch size 1024 (ppo_stage1.py:28)."""
    import torch
    from rl_collision_avoidance_b200 import _lib
    from rl_collision_avoidance_b200.model.net import Adam, CNNPolicy, _ptr
    pol = CNNPolicy(device=str(dev), max_batch=max(n_agents, 1024), seed=0)
    opt = Adam(pol.parameters(), lr=5e-5)
    lib = pol.lib

    def timeit(fn, n=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n * 1e3

    out = {'tensor_cores': 'tcgen05 3xTF32: conv tower forward/backward + fc1 forward/dW/dX'}
    for nb, key in ((n_agents, 'policy_forward_us'), (1024, 'ppo_minibatch_step_us')):
        obs = torch.rand(nb, 1536, device=dev) - 0.5
        gs = torch.rand(nb, 4, device=dev)
        v, mean = torch.empty(nb, device=dev), torch.empty(nb, 2, device=dev)
        act, lp = torch.rand(nb, 2, device=dev), torch.rand(nb, device=dev) - 1
        adv, tgt, losses = torch.randn(nb, device=dev), torch.randn(nb, device=dev), torch.zeros(3, device=dev)
        ws, st = pol._workspace(nb), pol._stream()

        def fwd():
            _lib.check(lib.rlca_policy_forward(ws, _ptr(pol.flat), _ptr(obs), _ptr(gs), nb, _ptr(v), _ptr(mean), st))

        def step():
            fwd()
            _lib.check(lib.rlca_ppo_loss_fwd_bwd(ws, _ptr(pol.flat), _ptr(v), _ptr(mean), _ptr(act), _ptr(lp), _ptr(adv),
                                                 _ptr(tgt), nb, 0.1, 5e-4, 20.0, _ptr(losses), st))
            _lib.check(lib.rlca_policy_backward(ws, _ptr(pol.flat), _ptr(obs), _ptr(gs), nb, _ptr(pol.grad), st))
            opt.step()

        out[key] = {'batch': nb, 'us': timeit(fwd if key == 'policy_forward_us' else step)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--e2e-steps', type=int, default=200)
    ap.add_argument('--cpu-steps', type=int, default=40)
    ap.add_argument('--e2e-sweep', action='store_true', help='also time step_host for several host-chunk counts')
    ap.add_argument('--ctas-per-world', type=int, default=0)
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    import __graft_entry__ as g
    if rank == 0:
        g.build(quiet=True)

    if args.impl == 'reference':
        if rank != 0:
            return
        val, cores, sample, ms = run_cpu(args.steps, args.warmup, budget_s=90.0)
        line = {
            'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'agent-steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'stage1 rink arena, 171 worlds x 24 robots = 4104 agents, 512 beams '
                                   '(reference CPU path = oracle port of Stage semantics; Stage/ROS/mpi4py '
                                   'are not installable here)', 'seed': 0},
            'cpu_baseline': {'value': val, 'unit': 'agent-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
        }
        print(json.dumps(line), flush=True)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from helpers import random_actions
    from rl_collision_avoidance_b200.stage_world import StageWorld

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world_size > 1:
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
    if rank != 0:
        g.build(quiet=True)

    env = StageWorld(BEAMS, index=0, scenario='stage1', num_worlds=WORLDS_PER_GPU, device=dev, seed=0,
                     auto_reset=True, world_offset=rank * WORLDS_PER_GPU, ctas_per_world=args.ctas_per_world)
    env.reset_pose()
    N = env.N
    rng = np.random.default_rng(1000 + rank)
    acts = [torch.from_numpy(random_actions(rng, N)).to(dev) for _ in range(64)]
    # rollout-style obs ring: 128 slots x N x 512 f32 = 1.08 GB > 126 MB L2, so consecutive
    # ticks never re-hit obs lines in L2 (the 0.26 MB simulator state is L2-resident by design)
    ring = torch.empty(128, N, BEAMS, device=dev)

    def tick(i):
        env.control_vel(acts[i % 64], obs_out=ring[i % 128])

    def sync():
        torch.cuda.synchronize(dev)
        if world_size > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        tick(i)
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = env.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        tick(i)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = env.launch_count - l0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the host-buffer call (pinned action H2D, tick, obs/reward/flags/gs D2H, sync inside every call)
    a_host = [torch.from_numpy(random_actions(rng, N)).pin_memory() for _ in range(8)]

    def time_e2e(want_obs=True, chunks=0, mode=1):
        env.set_host_chunks(chunks)
        env.set_host_zero_copy(mode)
        for i in range(5):
            env.step_host(a_host[i % 8], want_obs=want_obs)
        sync()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            env.step_host(a_host[i % 8], want_obs=want_obs)
        torch.cuda.synchronize(dev)
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world_size > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        env.set_host_chunks(0)
        env.set_host_zero_copy(1)
        return N * world_size * args.e2e_steps / float(te.item())

    e2e_val = time_e2e()                         # library default: no DMA, the kernel reads/writes pinned host memory
    e2e_serial = time_e2e(chunks=1, mode=0)      # copy in, one launch, copy out (the earlier round-1 number, for comparison)
    # variant for callers that keep the policy on the device: same call, but the observations stay in HBM
    # (action H2D + tick + reward/flags/goal-speed D2H + sync).  Reported next to `e2e`, never instead of it.
    e2e_noobs = time_e2e(want_obs=False)
    e2e_sweep = None
    if args.e2e_sweep:
        e2e_sweep = {f'mode{m}_chunks{k}': time_e2e(chunks=k, mode=m) for m in (0, 2) for k in (1, 2, 3, 4, 8)}

    if rank == 0:
        value = N * world_size * args.steps / (ms_max * 1e-3)
        per_launch_s = ms_max * 1e-3 / args.steps
        peak, peak_src = measured_peak()
        ach = N * alg_bytes(BEAMS) / per_launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json')
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic = json.load(f).get('dram_bytes_per_launch')
            except Exception:
                traffic = None
        line = {
            'metric': METRIC, 'value': value, 'unit': 'agent-steps/s', 'n_gpus': world_size, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'stage1 rink arena (100x100 cells @0.2 m), {WORLDS_PER_GPU} worlds x {ROBOTS} '
                                   f'robots = {N} agents per GPU, {BEAMS} beams, fov pi, range 6 m, dt 0.1 s, '
                                   'v~U[0,1] w~U[-1,1], auto-reset on done, seed 0',
                       'agents_per_gpu': N, 'beams': BEAMS, 'parallelism': f'worlds sharded over {world_size} GPU(s), '
                       'no data-path collective',
                       'l2': 'obs written round-robin into a 128-slot rollout ring (1.08 GB > 126 MB L2)',
                       'ctas_per_world': args.ctas_per_world or 'auto'},
            'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': N * alg_bytes(BEAMS),
                         'note': 'per-GPU; the march is issue/shared-memory bound, not HBM bound (DESIGN.md §6)'},
            'e2e': {'value': e2e_val, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': N * 8 * world_size,
                    'd2h_bytes_per_step': N * (4 * BEAMS + 4 + 4 + 16) * world_size, 'steps': args.e2e_steps,
                    'how': 'rlca_env_step_host with pinned host buffers, library default: the tick kernel reads the '
                           'actions from host memory and mirrors obs/reward/flags/gs to it over PCIe while it runs (no '
                           'DMA operations); every call ends with a stream synchronize',
                    'serial_value': e2e_serial},
            'e2e_obs_on_device': {'value': e2e_noobs, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': N * 8 * world_size,
                                  'd2h_bytes_per_step': N * (4 + 4 + 16) * world_size,
                                  'note': 'same host-buffer call with the scans left in HBM for an on-device policy'},
            'gpu_launches': int(launches),
            'clocks': clocks,
        }
        if e2e_sweep is not None:
            line['e2e_sweep_host_chunks'] = e2e_sweep
        if world_size == 1:
            line['learner'] = time_learner(dev, N)
        if world_size == 1 and not args.no_cpu:
            val, cores, sample, _ = run_cpu(args.cpu_steps, 3, budget_s=25.0)
            line['cpu_baseline'] = {'value': val, 'unit': 'agent-steps/s', 'cores': cores, 'kind': 'port',
                                    'sample': sample}
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
