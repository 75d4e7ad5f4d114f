"""Scenario definitions: constants of the reference's three environments and the
spawn/goal tables, as data for the simulator kernels.

Sources (under /root/reference): stage_world1.py, stage_world2.py,
circle_world.py (rewards, thresholds, timeouts), worlds/*.world (geometry),
model/utils.py:6-63 (tables; shipped here as assets/scenarios.json produced by
tools/build_assets.py).  SURVEY.md Appendix B lists every constant.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import numpy as np

from .worldfile import WorldMap, load_map

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets')


@dataclass
class Scenario:
    name: str
    scenario_id: int           # 0 stage1, 1 stage2, 2 circle
    robots_per_world: int
    timeout: int
    w_threshold: float
    pre_distance_zero: int
    map: WorldMap
    init_tab: np.ndarray       # (R,4) x,y,theta,random_flag  float32
    goal_tab: np.ndarray       # (R,4) gx,gy,random_flag,group id   float32
    groups: tuple = ()         # stage-2 group boundaries (model/utils.py:83)


# constants shared by the three env clients (stage_world1.py:34,183-204)
COMMON = dict(dt=0.1, range_max=6.0, fov=math.pi, raw_beams=512, half_len=0.22, half_wid=0.19,
              goal_radius=0.5, reward_arrive=15.0, reward_collision=-15.0, progress_gain=2.5,
              w_penalty=-0.1, v_min=0.0, v_max=1.0, w_min=-1.0, w_max=1.0)


def _tables():
    with open(os.path.join(ASSETS, 'scenarios.json')) as f:
        return json.load(f)


def make_scenario(name: str, map_: WorldMap | None = None, robots_per_world: int | None = None) -> Scenario:
    name = name.lower()
    if name == 'stage1':
        m = map_ or load_map(os.path.join(ASSETS, 'stage1_map.npz'))
        R = robots_per_world or 24
        init = np.zeros((R, 4), np.float32)
        k = min(R, len(m.init_poses))
        init[:k, :3] = m.init_poses[:k]
        goal = np.zeros((R, 4), np.float32)
        return Scenario('stage1', 0, R, 150, 1.05, 0, m, init, goal)
    if name == 'stage2':
        m = map_ or load_map(os.path.join(ASSETS, 'stage2_map.npz'))
        t = _tables()['stage2']
        R = 44
        init = np.zeros((R, 4), np.float32)
        init[:, :3] = np.asarray(t['init_pose'], np.float64)
        goal = np.zeros((R, 4), np.float32)
        g = np.asarray(t['goal_point'], np.float64)          # 34 rows (model/utils.py:55-63)
        goal[:len(g), :2] = g
        init[34:44, 3] = 1.0                                  # stage_world2.py:211 random spawn
        goal[34:44, 2] = 1.0                                  # stage_world2.py:165 random goal
        grp = t['groups']                                     # model/utils.py:83 group boundaries
        for gi, (a, b) in enumerate(zip(grp[:-1], grp[1:])):
            goal[a:b, 3] = gi
        return Scenario('stage2', 1, R, 200, 1.05, 1, m, init, goal, groups=tuple(t['groups']))
    if name == 'circle':
        m = map_ or load_map(os.path.join(ASSETS, 'circle_map.npz'))
        t = _tables()['circle']
        R = 50
        init = np.zeros((R, 4), np.float32)
        init[:, :3] = np.asarray(t['init_pose'], np.float64)
        goal = np.zeros((R, 4), np.float32)
        goal[:, :2] = np.asarray(t['goal_point'], np.float64)
        return Scenario('circle', 2, R, 10000, 0.7, 1, m, init, goal)
    raise ValueError(f'unknown scenario {name!r}')


def fill_config(cfg, sc: Scenario, *, num_worlds: int, beams: int, raw_beams: int | None = None,
                auto_reset: bool = False, seed: int = 0, world_offset: int = 0, max_reject: int = 4096):
    """Fill a ctypes config structure (EnvConfig or the oracle's mirror) from a Scenario."""
    import ctypes as C
    f32 = lambda v: C.c_float(v).value
    m = sc.map
    cfg.robots_per_world = sc.robots_per_world
    cfg.num_worlds = num_worlds
    cfg.beams = beams
    cfg.raw_beams = raw_beams if raw_beams is not None else max(COMMON['raw_beams'], beams)
    cfg.grid_w, cfg.grid_h = m.grid_w, m.grid_h
    cfg.origin_cx, cfg.origin_cy = m.origin_cx, m.origin_cy
    cfg.resolution = m.resolution
    cfg.ppm = f32(1.0 / m.resolution)
    cfg.dt = COMMON['dt']
    cfg.inv_dt = np.float32(1.0) / np.float32(COMMON['dt'])
    cfg.range_max = COMMON['range_max']
    cfg.range_cells = np.float32(cfg.ppm) * np.float32(COMMON['range_max'])
    cfg.fov = COMMON['fov']
    for k in ('half_len', 'half_wid', 'goal_radius', 'reward_arrive', 'reward_collision', 'progress_gain',
              'w_penalty', 'v_min', 'v_max', 'w_min', 'w_max'):
        setattr(cfg, k, COMMON[k])
    cfg.w_threshold = sc.w_threshold
    cfg.timeout = sc.timeout
    cfg.pre_distance_zero = sc.pre_distance_zero
    cfg.scenario = sc.scenario_id
    cfg.auto_reset = int(auto_reset)
    cfg.max_reject = max_reject
    cfg.world_offset = world_offset
    cfg.seed = seed
    return cfg
