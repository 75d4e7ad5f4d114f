"""Pins of the CPU oracle against everything the reference's own tests assert for this path.

The reference's only tests of the simulator are four rostests and a rate test that need a live
stageros (stage_ros-add_pose_and_crash/test/cmdpose_tests.py:87-203, test/hztest.xml:8-42).
They pin qualitative facts; each is restated here against the oracle.  Nothing in the reference
pins lidar ranges / collision flags / rewards numerically ("parity unpinned"), so
those get known-answer tests built from first principles in test_oracle_raycast.py."""
import math

import numpy as np
import pytest

from oracle.oracle import OracleWorld, OrcConfig, sincosf
from rl_collision_avoidance_b200.scenarios import Scenario, fill_config
from rl_collision_avoidance_b200.worldfile import WorldMap


def empty_world(R=2, cells=400, res=0.2, scenario_id=2, init=None, goals=None, **kw):
    m = WorldMap(cells=np.zeros((cells, cells), np.uint8), resolution=res, origin_cx=cells // 2, origin_cy=cells // 2,
                 init_poses=np.zeros((R, 3)))
    init_tab = np.zeros((R, 4), np.float32)
    goal_tab = np.zeros((R, 4), np.float32)
    if init is not None:
        init_tab[:, :3] = init
    if goals is not None:
        goal_tab[:, :2] = goals
    sc = Scenario('unit', scenario_id, R, kw.pop('timeout', 10000), kw.pop('w_threshold', 1.05), 0, m, init_tab, goal_tab)
    cfg = fill_config(OrcConfig(), sc, num_worlds=1, beams=kw.pop('beams', 16), raw_beams=kw.pop('raw_beams', None),
                      auto_reset=False, seed=0)
    return sc, OracleWorld(cfg, m.cells, init_tab, goal_tab)


def test_cmdvel_x_moves_along_heading_only():
    # cmdpose_tests.py:87-108 test_cmdvel_x: linear.x = 1.0 for 3 s changes x only
    sc, w = empty_world(init=[[0, 0, 0], [10, 10, 0]], goals=[[30, 0], [30, 10]])
    w.reset_world()
    w.reset_pose()
    a = np.array([[1.0, 0.0], [0.0, 0.0]], np.float32)
    for _ in range(30):
        w.step(a)
    assert abs(w.pose[0, 0] - 3.0) < 1e-5          # 30 ticks x 0.1 s x 1 m/s  (hztest.xml: 10 Hz -> dt 0.1)
    assert w.pose[0, 1] == 0.0 and w.pose[0, 2] == 0.0
    assert np.array_equal(w.pose[1, :3], np.array([10, 10, 0], np.float32))   # zero command: untouched


def test_cmdvel_yaw_changes_only_yaw():
    # cmdpose_tests.py:112-133 test_cmdvel_yaw: angular.z = 0.25 changes only the heading
    sc, w = empty_world(init=[[1.5, -2.5, 0.3], [10, 10, 0]], goals=[[30, 0], [30, 10]])
    w.reset_world()
    w.reset_pose()
    a = np.array([[0.0, 0.25], [0.0, 0.0]], np.float32)
    for _ in range(30):
        w.step(a)
    assert w.pose[0, 0] == np.float32(1.5) and w.pose[0, 1] == np.float32(-2.5)
    assert abs(w.pose[0, 2] - (0.3 + 0.75)) < 1e-5


def test_teleport_sets_pose_exactly_and_keeps_stall():
    # cmdpose_tests.py:136-203 test_pose / test_pose_stamped: (x, y, yaw) land exactly
    sc, w = empty_world(init=[[42.0, -42.0, 0.9], [0, 0, 0]], goals=[[0, 0], [1, 1]], cells=600)
    w.reset_world()
    w.meta[0, 2] = 1                       # pretend it was stalled
    w.reset_pose()
    assert np.array_equal(w.pose[0, :3], np.array([42.0, -42.0, 0.9], np.float32))
    assert w.meta[0, 2] == 1               # SetPose does not touch the stall flag (stageros.cpp:282-296)
    w.reset_world()
    assert w.meta[0, 2] == 0               # reset_positions clears it (stageros.cpp:266)


def test_heading_wraps_to_minus_pi_pi():
    sc, w = empty_world(init=[[0, 0, 3.1], [10, 10, 0]], goals=[[30, 0], [30, 10]])
    w.reset_world()
    w.reset_pose()
    a = np.array([[0.0, 1.0], [0.0, 0.0]], np.float32)
    for _ in range(10):
        w.step(a)
    assert -math.pi < w.pose[0, 2] <= math.pi
    assert abs(w.pose[0, 2] - (3.1 + 1.0 - 2 * math.pi)) < 1e-5


def test_sincos_accuracy():
    x = np.linspace(-7, 7, 4001).astype(np.float32)
    s, c = sincosf(x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 2e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 2e-7


def test_reward_and_done_rules():
    # stage_world1.py:180-211 with hand-computable numbers
    sc, w = empty_world(R=3, scenario_id=2, init=[[0, 0, 0], [5, 5, 0], [-5, -5, 0]],
                        goals=[[0.55, 0], [9, 5], [-5, 0]], timeout=3)
    w.cfg.pre_distance_zero = 0
    w.reset_world()
    w.reset_pose()
    assert abs(w.pose[0, 3] - 0.55) < 1e-6           # pre_distance = true distance (stage_world1.py:174-177)
    a = np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 0.0]], np.float32)
    w.step(a)
    # robot 0: 0.55 -> 0.45 < 0.5  => reach goal, reward 15
    assert w.flags[0, 0] == 1 and w.flags[0, 2] == 1 and w.reward[0] == 15.0
    # robot 1: progress 0.1 m * 2.5
    assert w.flags[1, 0] == 0 and abs(w.reward[1] - 0.25) < 1e-5
    # robot 2 idles; times out when t > 3 (4th call)
    assert w.flags[2, 0] == 0 and w.reward[2] == 0.0
    for _ in range(3):
        w.step(a)
    assert w.flags[2, 0] == 1 and w.flags[2, 2] == 3


def test_rotation_penalty_only_in_circle_threshold():
    # |w| > 1.05 can never fire for a clipped action; circle's 0.7 can (circle_world.py:195-196)
    sc, w = empty_world(init=[[0, 0, 0], [10, 10, 0]], goals=[[30, 0], [30, 10]], w_threshold=0.7)
    w.reset_world()
    w.reset_pose()
    w.step(np.array([[0.0, 1.0], [0.0, 0.5]], np.float32))
    assert abs(w.reward[0] - (-0.1)) < 1e-5
    assert w.reward[1] == 0.0


def test_pre_distance_zero_quirk():
    # stage_world2.py:170 / circle_world.py:166: first-step reward is -2.5 * d0
    sc, w = empty_world(init=[[0, 0, 0], [10, 10, 0]], goals=[[4, 0], [30, 10]])
    w.cfg.pre_distance_zero = 1
    w.reset_world()
    w.reset_pose()
    w.step(np.array([[0.0, 0.0], [0.0, 0.0]], np.float32))
    assert abs(w.reward[0] - (-2.5 * 4.0)) < 1e-5


def test_robot_robot_collision_reverts_and_stalls():
    sc, w = empty_world(init=[[0, 0, 0], [0.7, 0, math.pi]], goals=[[30, 0], [-30, 0]])
    w.reset_world()
    w.reset_pose()
    a = np.array([[1.0, 0.0], [1.0, 0.0]], np.float32)
    crashed_at = None
    for t in range(6):
        before = w.pose.copy()
        w.step(a)
        if w.flags[:, 1].any():
            crashed_at = t
            hit = w.flags[:, 1] == 1
            assert np.array_equal(w.pose[hit, :3], before[hit, :3])      # pose restored
            assert np.all(w.reward[hit] <= -14.0)
            break
    assert crashed_at is not None


def test_fp32_tick_tracks_double_precision_stage_arithmetic():
    """Deviation D4: Stage integrates in double (x += v dt cos th, y += v dt sin th,
    th = normalize(th + w dt), old heading used for the translation) and the Python client computes reward and local
    goal in float64 (stage_world1.py:155-160,180-211).  The oracle's fp32 contract must stay within the north-star's
    1e-4 of that arithmetic over a whole stage-1 episode (150 ticks) of random commands."""
    rng = np.random.default_rng(42)
    R = 8
    init = np.column_stack([rng.uniform(-8, 8, R), rng.uniform(-8, 8, R), rng.uniform(-math.pi, math.pi, R)])
    goals = np.column_stack([rng.uniform(25, 30, R), rng.uniform(25, 30, R)])      # never reached
    init[:, :2] += np.arange(R)[:, None] * 40.0            # far apart: no robot-robot contact
    goals += np.arange(R)[:, None] * 40.0
    sc, w = empty_world(R=R, cells=3200, init=init, goals=goals, timeout=10000, w_threshold=0.7)
    w.reset_world()
    w.reset_pose()
    x = w.pose[:, 0].astype(np.float64)
    y = w.pose[:, 1].astype(np.float64)
    th = w.pose[:, 2].astype(np.float64)
    gx, gy = w.goal[:, 0].astype(np.float64), w.goal[:, 1].astype(np.float64)
    d_prev = np.hypot(gx - x, gy - y)                       # pre_distance = true distance (stage_world1.py:174-177)
    worst = dict(pos=0.0, pos0=0.0, th=0.0, rew=0.0, lg=0.0)
    for t in range(150):
        a = np.column_stack([rng.uniform(0, 1, R), rng.uniform(-1, 1, R)]).astype(np.float32)
        w.step(a)
        v, om = a[:, 0].astype(np.float64), a[:, 1].astype(np.float64)
        th_old = th.copy()
        x = x + v * 0.1 * np.cos(th)
        y = y + v * 0.1 * np.sin(th)
        th = th + om * 0.1
        th = np.where(th > math.pi, th - 2 * math.pi, np.where(th <= -math.pi, th + 2 * math.pi, th))
        w_gt = (np.mod(th - th_old + math.pi, 2 * math.pi) - math.pi) / 0.1       # stageros.cpp:581-593
        d = np.hypot(gx - x, gy - y)
        rew = 2.5 * (d_prev - d) + np.where(np.abs(w_gt) > 0.7, -0.1 * np.abs(w_gt), 0.0)
        d_prev = d
        lgx = (gx - x) * np.cos(th) + (gy - y) * np.sin(th)                         # stage_world1.py:155-160
        lgy = -(gx - x) * np.sin(th) + (gy - y) * np.cos(th)
        worst['pos'] = max(worst['pos'], np.abs(w.pose[:, 0] - x).max(), np.abs(w.pose[:, 1] - y).max())
        worst['pos0'] = max(worst['pos0'], abs(w.pose[0, 0] - x[0]), abs(w.pose[0, 1] - y[0]))
        worst['th'] = max(worst['th'], np.abs(w.pose[:, 2] - th).max())
        # |w_gt| near the 0.7 threshold may flip the penalty term between fp32 and double: compare away from it
        clear = np.abs(np.abs(w_gt) - 0.7) > 1e-4
        worst['rew'] = max(worst['rew'], np.abs(w.reward - rew)[clear].max() if clear.any() else 0.0)
        worst['lg'] = max(worst['lg'], np.abs(w.gs[:, 0] - lgx).max(), np.abs(w.gs[:, 1] - lgy).max())
    assert not w.flags[:, 0].any()
    # robot 0 moves at arena-scale coordinates (|x| < 25 m): the north-star's 1e-4.  The others sit up to 300 m out
    # (40 m apart so that nobody touches), where one fp32 ulp is already 3e-5: 1e-3.
    assert worst['pos0'] < 1e-4 and worst['th'] < 1e-4, worst
    assert worst['pos'] < 1e-3 and worst['rew'] < 2e-3 and worst['lg'] < 2e-3, worst
