"""The CPU oracle against golden vectors produced by the reference's OWN env-client code (no GPU).

tools/make_golden_env.py imports /root/reference/stage_world1.py, stage_world2.py and circle_world.py unmodified (ROS
imports replaced by empty stand-ins) and calls their get_laser_observation / get_local_goal / get_reward_and_terminate /
generate_goal_point / generate_random_pose methods on a bare object; the outputs are committed as
tests/golden/env_golden.npz.  This pins the Python side of the environment (observation map,
local goal, reward / done incl. the result-string precedence, goal bookkeeping, spawn regions) to the reference's code.
The Stage side (integration, collision, raytrace) stays unpinned: libstage is absent."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from oracle.oracle import OrcConfig, load
from test_oracle_pins import empty_world

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'env_golden.npz')
SCEN = {'stage1': dict(scenario_id=0, timeout=150, w_threshold=1.05, pre_zero=0),
        'stage2': dict(scenario_id=1, timeout=200, w_threshold=1.05, pre_zero=1),
        'circle': dict(scenario_id=2, timeout=10000, w_threshold=0.7, pre_zero=1)}


@pytest.mark.parametrize('nb', [512, 360, 180])
def test_observation_map_matches_reference_get_laser_observation(built, nb):
    """stage_world1.py:122-140 (identical in the other two clients): NaN / inf -> 6.0, symmetric nearest-index
    sub-sampling of the 512 raw ranges, scan / 6 - 0.5.  The oracle's beam table must pick the same raw beams."""
    g = np.load(GOLD)
    lib = load()
    cfg = OrcConfig()
    cfg.beams, cfg.raw_beams, cfg.fov = nb, 512, math.pi
    cb, sb = np.zeros(nb, np.float32), np.zeros(nb, np.float32)
    idx = np.zeros(nb, np.int32)
    lib.orc_beam_table(C.byref(cfg), cb.ctypes.data_as(C.c_void_p), sb.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p))
    scan = g['scan_raw'].copy()
    scan[~np.isfinite(scan)] = 6.0                     # the simulator reports range_max for a miss; the client maps NaN / inf
    ours = scan[idx] / 6.0 - 0.5
    for name in SCEN:
        assert np.allclose(ours, g[f'obs_{name}_{nb}'], atol=1e-12), name
    # and the normalisation as the oracle computes it in fp32: fmaf(range, 1/6, -0.5)
    f32 = np.float32(scan[idx].astype(np.float32) * np.float32(1.0 / 6.0) - np.float32(0.5))
    assert np.abs(f32 - g[f'obs_stage1_{nb}']).max() < 1e-6


@pytest.mark.parametrize('name', ['stage1', 'stage2', 'circle'])
def test_local_goal_matches_reference(built, name):
    """get_local_goal (stage_world1.py:155-160): the goal rotated into the robot frame."""
    g = np.load(GOLD)
    state, goals = g['lg_state'], g['lg_goal']
    K = len(state)
    sc, w = empty_world(R=K, scenario_id=2, init=state, goals=goals)
    w.reset_world()
    w.pose[:, :3] = state.astype(np.float32)
    w.goal[:, :2] = goals.astype(np.float32)
    w.observe()
    assert np.abs(w.gs[:, :2] - g[f'local_goal_{name}']).max() < 2e-5           # fp32 vs the reference's float64


@pytest.mark.parametrize('name', ['stage1', 'stage2', 'circle'])
def test_reward_and_terminate_match_reference(built, name):
    """get_reward_and_terminate of each client (stage_world1.py:180-211, stage_world2.py:175-208,
    circle_world.py:171-203) on 50 cases incl. reach-goal + crash in one step, the |w| penalty thresholds (1.05 vs
    0.7), the three time-outs and the precedence of the result strings."""
    g = np.load(GOLD)
    cases, ref = g['reward_cases'], g[f'reward_{name}']
    K = len(cases)
    s = SCEN[name]
    init = np.zeros((K, 3))
    init[:, 0] = cases[:, 0] + 12.0 * (np.arange(K) % 8) - 40.0   # keep the robots apart: only their own state matters
    init[:, 1] = cases[:, 1] + 12.0 * (np.arange(K) // 8) - 40.0
    init[:, 2] = 0.3
    goal = cases[:, 2:4] + (init[:, :2] - cases[:, :2])
    sc, w = empty_world(R=K, cells=1200, scenario_id=2, init=init, goals=goal, timeout=s['timeout'],
                        w_threshold=s['w_threshold'])
    w.reset_world()
    w.pose[:, :3] = init.astype(np.float32)
    w.pose[:, 3] = cases[:, 4].astype(np.float32)              # self.distance before the call (= pre_distance of the step)
    w.goal[:, :2] = goal.astype(np.float32)
    w.meta[:, 0] = cases[:, 7].astype(np.int32)                # t
    w.meta[:, 2] = cases[:, 5].astype(np.int32)                # is_crashed (a robot that is not commanded keeps it)
    a = np.zeros((K, 2), np.float32)
    a[:, 1] = cases[:, 6]                                      # rotate on the spot: ground-truth w = the command
    assert np.all((cases[:, 5] == 0) | (cases[:, 6] == 0))     # crashed cases stand still (the flag is then untouched)
    w.step(a)
    assert np.abs(w.reward - ref[:, 0]).max() < 2e-5, np.argmax(np.abs(w.reward - ref[:, 0]))
    assert np.array_equal(w.flags[:, 0], ref[:, 1].astype(np.uint8))            # terminate
    assert np.array_equal(w.flags[:, 2], ref[:, 2].astype(np.uint8))            # result string code
    assert np.abs(w.pose[:, 3] - ref[:, 3]).max() < 2e-5                         # self.distance after the call


def test_generate_goal_point_bookkeeping_matches_reference(built):
    """Stage 1 sets pre_distance = distance to the new goal (stage_world1.py:171-177); stage 2 / circle set 0
    (stage_world2.py:170, circle_world.py:166) - the first-step reward quirk.  Table goals are the reference's."""
    g = np.load(GOLD)
    p1 = g['goal_stage1_pre']
    d = np.hypot(g['goal_stage1'][:, 2] - g['goal_stage1'][:, 0], g['goal_stage1'][:, 3] - g['goal_stage1'][:, 1])
    assert np.allclose(p1[:, 0], d) and np.allclose(p1[:, 1], d)
    assert np.all(g['goal_stage2_random'][:, 2:] == 0) and g['goal_stage2_table5'][2] == 0 and g['goal_circle_table7'][2] == 0
    from rl_collision_avoidance_b200.scenarios import make_scenario
    assert np.allclose(make_scenario('stage2').goal_tab[5, :2], g['goal_stage2_table5'][:2])
    assert np.allclose(make_scenario('circle').goal_tab[7, :2], g['goal_circle_table7'][:2])
    assert make_scenario('stage1').pre_distance_zero == 0 and make_scenario('stage2').pre_distance_zero == 1 \
        and make_scenario('circle').pre_distance_zero == 1


def test_spawn_regions_match_reference_rejection_loops(built):
    """generate_random_pose / generate_random_goal: the reference's own draws (np.random) and the oracle's (Philox) are
    different streams, so the comparison is of the accepted REGIONS: every draw of either side satisfies the
    reference's loop-exit condition, and both fill the region (stage_world1.py:251-274, stage_world2.py:250-287)."""
    g = np.load(GOLD)
    ref_pose, ref_goal = g['pose_stage1'], g['goal_stage1']
    assert np.all(np.hypot(ref_pose[:, 0], ref_pose[:, 1]) <= 9.0)
    dg = np.hypot(ref_goal[:, 2] - ref_goal[:, 0], ref_goal[:, 3] - ref_goal[:, 1])
    assert np.all(np.hypot(ref_goal[:, 2], ref_goal[:, 3]) <= 9.0) and np.all((dg >= 8.0) & (dg <= 10.0))
    # the oracle's stage-1 draws: same conditions, same spread
    from helpers import make_pair
    _, _, orc = make_pair('stage1', num_worlds=20, gpu=False, seed=3)
    orc.reset_world()
    orc.reset_pose()
    assert np.all(np.hypot(orc.pose[:, 0], orc.pose[:, 1]) <= 9.0 + 1e-5)
    d = np.hypot(orc.goal[:, 0] - orc.pose[:, 0], orc.goal[:, 1] - orc.pose[:, 1])
    assert np.all(np.hypot(orc.goal[:, 0], orc.goal[:, 1]) <= 9.0 + 1e-5) and np.all((d >= 8.0 - 1e-4) & (d <= 10.0 + 1e-4))
    assert abs(np.hypot(orc.pose[:, 0], orc.pose[:, 1]).mean() - np.hypot(ref_pose[:, 0], ref_pose[:, 1]).mean()) < 0.5
    assert np.all((orc.pose[:, 2] > -math.pi - 1e-6) & (orc.pose[:, 2] <= math.pi + 1e-6))
    assert np.all((ref_pose[:, 2] >= 0) & (ref_pose[:, 2] <= 2 * math.pi))       # the reference draws [0, 2 pi); Stage wraps it
    # stage 2, robots 34..43: x in [9, 19], y in [-19, -9] u [-5, -1]... as the reference's loop leaves them
    rp, rg = g['pose_stage2_random'], g['goal_stage2_random']
    _, _, o2 = make_pair('stage2', num_worlds=30, gpu=False, seed=4)
    o2.reset_world()
    o2.reset_pose()
    mine = o2.pose.reshape(30, 44, 4)[:, 34:44, :2].reshape(-1, 2)
    for arr in (rp[:, :2], mine):
        assert arr[:, 0].min() >= 9.0 - 1e-5 and arr[:, 0].max() <= 19.0 + 1e-5
        assert arr[:, 1].min() >= -19.0 - 1e-5 and arr[:, 1].max() <= -1.0 + 1e-5
        assert not np.any((arr[:, 1] > -9.0 + 1e-5) & (arr[:, 1] < -5.0 - 1e-5))          # the gap between the two bands
    # the rejection test is relative to where the robot stands: >= 7 m from (12, -5) for the reference's draws (the pose
    # tools/make_golden_env.py gave it), >= 7 m from the robot's world-file pose for the oracle's first reset_pose
    assert np.all(np.hypot(rp[:, 0] - 12.0, rp[:, 1] + 5.0) >= 7.0)
    init = np.tile(o2.init_tab[34:44, :2], (30, 1))
    assert np.all(np.hypot(mine[:, 0] - init[:, 0], mine[:, 1] - init[:, 1]) >= 7.0 - 1e-4)
    # and the goal of those robots is >= 7 m from the pose just drawn (stage_world2.py:270-287)
    mg = o2.goal.reshape(30, 44, 4)[:, 34:44, :2].reshape(-1, 2)
    assert np.all(np.hypot(mg[:, 0] - mine[:, 0], mg[:, 1] - mine[:, 1]) >= 7.0 - 1e-4)
    assert np.all(np.hypot(rg[:, 0] - 12.0, rg[:, 1] + 5.0) >= 7.0)
    assert rg[:, 0].min() >= 9.0 and rg[:, 0].max() <= 19.0
