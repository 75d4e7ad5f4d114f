"""ctypes wrapper of the CPU oracle (oracle/sim_oracle.c -> liboracle_sim.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs — never by the product
package.  Parity status of the simulator half: "parity unpinned" (libstage is
absent; see the header of sim_oracle.c and DESIGN.md §3).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liboracle_sim.so')


class OrcConfig(C.Structure):
    """orc_config of sim_oracle.c (declared independently of the product's struct)."""
    _fields_ = [
        ('robots_per_world', C.c_int32), ('num_worlds', C.c_int32),
        ('beams', C.c_int32), ('raw_beams', C.c_int32),
        ('grid_w', C.c_int32), ('grid_h', C.c_int32),
        ('origin_cx', C.c_int32), ('origin_cy', C.c_int32),
        ('resolution', C.c_float), ('ppm', C.c_float),
        ('dt', C.c_float), ('inv_dt', C.c_float),
        ('range_max', C.c_float), ('range_cells', C.c_float),
        ('fov', C.c_float),
        ('half_len', C.c_float), ('half_wid', C.c_float),
        ('goal_radius', C.c_float), ('reward_arrive', C.c_float),
        ('reward_collision', C.c_float), ('progress_gain', C.c_float),
        ('w_threshold', C.c_float), ('w_penalty', C.c_float),
        ('v_min', C.c_float), ('v_max', C.c_float), ('w_min', C.c_float), ('w_max', C.c_float),
        ('timeout', C.c_int32), ('pre_distance_zero', C.c_int32),
        ('scenario', C.c_int32), ('auto_reset', C.c_int32),
        ('max_reject', C.c_int32), ('world_offset', C.c_int32),
        ('seed', C.c_uint64),
    ]


_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


class OracleWorld:
    """Host-side twin of StageWorld's state: numpy arrays with the same layouts."""

    def __init__(self, cfg: OrcConfig, static_cells: np.ndarray, init_tab: np.ndarray, goal_tab: np.ndarray):
        self.lib = load()
        self.cfg = cfg
        self.static = np.ascontiguousarray(np.where(static_cells != 0, 254, 0).astype(np.uint8))
        assert self.static.shape == (cfg.grid_h, cfg.grid_w)
        self.init_tab = np.ascontiguousarray(init_tab, np.float32)
        self.goal_tab = np.ascontiguousarray(goal_tab, np.float32)
        N = cfg.robots_per_world * cfg.num_worlds
        self.N = N
        self.pose = np.zeros((N, 4), np.float32)
        self.goal = np.zeros((N, 4), np.float32)
        self.acc = np.zeros((N, 4), np.float32)
        self.meta = np.zeros((N, 4), np.int32)
        self.obs = np.zeros((N, cfg.beams), np.float32)
        self.reward = np.zeros(N, np.float32)
        self.flags = np.zeros((N, 4), np.uint8)
        self.gs = np.zeros((N, 4), np.float32)
        self.eplog = np.zeros((N, 8), np.float32)

    def reset_world(self):
        mask = np.zeros(self.N, np.uint8)
        self.lib.orc_reset(C.byref(self.cfg), _p(self.init_tab), _p(self.goal_tab), _p(mask), 1,
                           _p(self.pose), _p(self.goal), _p(self.acc), _p(self.meta))

    def reset_pose(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.lib.orc_reset(C.byref(self.cfg), _p(self.init_tab), _p(self.goal_tab), _p(m), 0,
                           _p(self.pose), _p(self.goal), _p(self.acc), _p(self.meta))
        self.observe()

    def generate_goal_point(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.lib.orc_reset(C.byref(self.cfg), _p(self.init_tab), _p(self.goal_tab), _p(m), 2,
                           _p(self.pose), _p(self.goal), _p(self.acc), _p(self.meta))
        self.observe()

    def observe(self):
        self.lib.orc_observe(C.byref(self.cfg), _p(self.static), _p(self.pose), _p(self.goal), _p(self.obs), _p(self.gs))

    def step(self, action: np.ndarray, live=None):
        a = np.ascontiguousarray(action, np.float32)
        lv = None if live is None else np.ascontiguousarray(live, np.uint8)
        self.lib.orc_step(C.byref(self.cfg), _p(self.static), _p(self.init_tab), _p(self.goal_tab), _p(a), _p(lv),
                          _p(self.pose), _p(self.goal), _p(self.acc), _p(self.meta),
                          _p(self.obs), _p(self.reward), _p(self.flags), _p(self.gs), _p(self.eplog))

    def raycast(self, pose: np.ndarray, normalise=False):
        pose = np.ascontiguousarray(pose, np.float32)
        out = np.zeros((self.N, self.cfg.beams), np.float32)
        self.lib.orc_raycast(C.byref(self.cfg), _p(self.static), _p(pose), _p(out), int(normalise))
        return out


def sincosf(x: np.ndarray):
    lib = load()
    x = np.ascontiguousarray(x, np.float32)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    fs, fc = C.c_float(), C.c_float()
    for i, v in enumerate(x.ravel()):
        lib.orc_sincosf(C.c_float(float(v)), C.byref(fs), C.byref(fc))
        s.ravel()[i] = fs.value
        c.ravel()[i] = fc.value
    return s, c


def num_threads():
    return int(load().orc_num_threads())


def set_threads(n: int):
    load().orc_set_threads(int(n))
