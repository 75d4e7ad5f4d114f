"""GPU parity at the sizes bench.py times (BASELINE.json configs 1-5), against the CPU oracle, bit for bit.

The small-size tests in test_env_parity_gpu.py sweep the code paths; these close the perimeter: every world of the
headline batch for longer than a stage-1 episode (time-outs at 150 ticks must occur), the stage-2 batch at its bench
size with the in-kernel group barrier, circle worlds driven until robots actually touch, and the stand-alone raycast
at the 65 536-robot sweep size for every beam count of config 5.  The oracle runs OpenMP over worlds, so each
comparison costs seconds."""
import numpy as np
import pytest
import torch

from helpers import make_pair, random_actions

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def _compare_tick(env, orc, t, every_obs=True):
    torch.cuda.synchronize()
    st = env.state
    for k, ref in (('pose', orc.pose), ('goal', orc.goal), ('acc', orc.acc), ('meta', orc.meta)):
        got = st[k].cpu().numpy()
        if not _eq(got, ref):
            rows = np.unique(np.nonzero(got.view(np.uint32) != ref.view(np.uint32))[0])
            raise AssertionError(f'tick {t}: state {k} differs at agents {rows[:8]} ({len(rows)} rows)')
    assert np.array_equal(env.flags.cpu().numpy(), orc.flags), f'tick {t}: flags'
    assert _eq(env.reward.cpu().numpy(), orc.reward), f'tick {t}: reward'
    assert _eq(env.gs.cpu().numpy(), orc.gs), f'tick {t}: gs'
    if every_obs:
        o = env.obs.cpu().numpy()
        if not _eq(o, orc.obs):
            bad = np.argwhere(o != orc.obs)
            raise AssertionError(f'tick {t}: obs differ at {bad[:4]} ({len(bad)} beams), max abs {np.abs(o - orc.obs).max()}')
    done = orc.flags[:, 0] != 0
    assert np.array_equal(env.eplog.cpu().numpy()[done], orc.eplog[done]), f'tick {t}: eplog'


def test_headline_all_171_worlds_220_ticks(built):
    """Config 1/headline: 171 stage-1 worlds x 24 robots x 512 beams, every world, 220 ticks (> the 150-tick
    time-out), state + flags + reward + scans compared every tick."""
    sc, env, orc = make_pair('stage1', num_worlds=171, beams=512, auto_reset=True, seed=0)
    env.reset_pose()
    orc.reset_world()
    orc.reset_pose()
    rng = np.random.default_rng(17)
    seen = np.zeros(4, np.int64)
    for t in range(220):
        # slow, gently turning robots in a third of the worlds survive to the time-out; the rest crash / arrive / re-spawn
        a = random_actions(rng, orc.N, wide=(t % 5 == 0))
        a[: orc.N // 3, 0] *= 0.1
        a[: orc.N // 3, 1] *= 0.2
        env.control_vel(torch.from_numpy(a).cuda())
        orc.step(a)
        _compare_tick(env, orc, t)
        seen += np.bincount(orc.flags[:, 2], minlength=4)[:4]
    assert seen[1] > 0 and seen[2] > 100 and seen[3] > 100, f'arrived/crashed/timed out = {seen[1:]}'


def test_stage2_94_worlds_250_ticks_group_barrier(built):
    """Config 3 at its bench size: 94 stage-2 worlds x 44 robots = 4136 agents, auto_reset=2 (finished robots idle until
    their group is done, then the group re-spawns inside the kernel), 250 ticks (> the 200-tick time-out)."""
    sc, env, orc = make_pair('stage2', num_worlds=94, beams=512, auto_reset=2, seed=5)
    env.reset_pose()
    orc.reset_world()
    orc.reset_pose()
    rng = np.random.default_rng(23)
    resets = idle = 0
    seen = np.zeros(4, np.int64)
    for t in range(250):
        a = random_actions(rng, orc.N)
        a[::3, 0] *= 0.05                      # a third crawl: their groups wait on them until the time-out
        env.control_vel(torch.from_numpy(a).cuda())
        orc.step(a)
        _compare_tick(env, orc, t, every_obs=(t % 3 == 0 or t > 240))
        resets += int(orc.flags[:, 3].sum())
        idle += int((orc.meta[:, 3] != 0).sum())
        seen += np.bincount(orc.flags[:, 2], minlength=4)[:4]
    assert resets > 1000 and idle > 1000 and seen[2] > 100 and seen[3] > 0, (resets, idle, seen)


def test_circle_8_worlds_300_ticks_with_contact(built):
    """Config 4's world: circle.world, 50 robots per world with antipodal goals, steered at their goals so that they
    meet in the middle: robot-robot collisions (crash flags far from any wall), re-spawns on the circle, the |w| > 0.7
    penalty — all bit-exact for 300 ticks on 8 worlds."""
    sc, env, orc = make_pair('circle', num_worlds=8, beams=512, auto_reset=1, seed=9)
    env.reset_pose()
    orc.reset_world()
    orc.reset_pose()
    rng = np.random.default_rng(31)
    crashes = arrived = 0
    seen_close = False
    for t in range(300):
        gx, gy = orc.gs[:, 0], orc.gs[:, 1]                      # goal in the robot frame (get_local_goal)
        a = np.empty((orc.N, 2), np.float32)
        a[:, 0] = rng.uniform(0.7, 1.0, orc.N)
        a[:, 1] = np.clip(2.0 * np.arctan2(gy, gx) + rng.normal(0, 0.15, orc.N), -1, 1)
        env.control_vel(torch.from_numpy(a).cuda())
        orc.step(a)
        _compare_tick(env, orc, t, every_obs=(t % 4 == 0 or t > 230))
        cr = orc.flags[:, 2] == 2
        crashes += int(cr.sum())
        arrived += int((orc.flags[:, 2] == 1).sum())
        seen_close |= bool((orc.obs < -0.4).any())                # a neighbour closer than 0.6 m in somebody's scan
    # the circle arena has no obstacles inside: every crash away from the boundary is robot against robot
    assert crashes > 20 and seen_close, (crashes, arrived)


@pytest.mark.parametrize('beams', [180, 360, 512, 1024])
def test_raycast_sweep_size_65544_robots(built, beams):
    """Config 5: rlca_raycast at 2731 stage-1 worlds x 24 = 65 544 robots for each beam count of the sweep, poses
    scattered over the arena (robots see each other and the walls), raw and normalised ranges vs orc_raycast."""
    worlds = 2731
    sc, env, orc = make_pair('stage1', num_worlds=worlds, beams=beams, raw_beams=max(512, beams), seed=1)
    rng = np.random.default_rng(100 + beams)
    pose = np.zeros((orc.N, 4), np.float32)
    pose[:, 0] = rng.uniform(-9.5, 9.5, orc.N)
    pose[:, 1] = rng.uniform(-9.5, 9.5, orc.N)
    pose[:, 2] = rng.uniform(-np.pi, np.pi, orc.N)
    pose[7, :2] = [40.0, 3.0]                                    # off the floor plan: every beam misses
    got = env.raycast(torch.from_numpy(pose).cuda(), normalise=False).cpu().numpy()
    ref = orc.raycast(pose, normalise=False)
    assert _eq(got, ref), f'{beams} beams: {np.count_nonzero(got != ref)} ranges differ, max abs {np.abs(got - ref).max()}'
    assert np.all(ref[7] == 6.0) and ref.min() >= 0.0 and ref.max() <= 6.0
    assert (ref < 6.0).mean() > 0.3                               # the sweep is not a trivial all-miss workload
    gotn = env.raycast(torch.from_numpy(pose).cuda(), normalise=True).cpu().numpy()
    assert _eq(gotn, orc.raycast(pose, normalise=True))
