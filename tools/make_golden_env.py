"""Golden vectors for the Python side of the environment from the reference's OWN code.

Run HERE (where /root/reference exists): python tools/make_golden_env.py
The reference's env clients (stage_world1.py, stage_world2.py, circle_world.py) import rospy / tf / ROS message
packages at module level, which do not exist here - but the arithmetic the hot path must reproduce lives in plain
methods: get_laser_observation (beam sub-sampling + scan/6 - 0.5), get_local_goal, get_reward_and_terminate,
generate_goal_point, generate_random_pose / generate_random_goal.  This script puts empty stand-in modules for the ROS
imports into sys.modules, imports the three reference files UNMODIFIED (read-only), and calls those methods on a bare
object that carries only the attributes they read.  Nothing of Stage runs: these vectors pin rows a8-a11 of SURVEY §8(a)
(observation map, local goal, reward / done, goal bookkeeping, spawn acceptance regions), not the simulator core.
Writes tests/golden/env_golden.npz; tests/test_oracle_reference_env.py compares the CPU oracle with it.
"""
import builtins
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('RLCA_REFERENCE', '/root/reference')
RESULT_CODE = {0: 0, 'Reach Goal': 1, 'Crashed': 2, 'Time out': 3}


def import_reference():
    builtins.xrange = range                                   # py2 builtin used by get_laser_observation
    for name in ('rospy', 'tf', 'geometry_msgs', 'geometry_msgs.msg', 'nav_msgs', 'nav_msgs.msg', 'sensor_msgs',
                 'sensor_msgs.msg', 'rosgraph_msgs', 'rosgraph_msgs.msg', 'std_srvs', 'std_srvs.srv', 'std_msgs',
                 'std_msgs.msg'):
        m = types.ModuleType(name)
        for attr in ('Twist', 'Pose', 'Odometry', 'LaserScan', 'Clock', 'Empty', 'Int8'):
            setattr(m, attr, type(attr, (), {}))
        sys.modules.setdefault(name, m)
    sys.modules['rospy'].is_shutdown = lambda: False
    sys.path.insert(0, REF)
    import circle_world
    import stage_world1
    import stage_world2
    return {'stage1': stage_world1.StageWorld, 'stage2': stage_world2.StageWorld, 'circle': circle_world.StageWorld}


class Bare:
    """Carries the attributes the reference methods read; the methods themselves are the reference's."""

    def __init__(self, cls, **kw):
        self._cls = cls
        self.goal_size = 0.5                                  # stage_world1.py:34
        self.__dict__.update(kw)

    def __getattr__(self, name):                              # get_self_stateGT etc.: the reference's own accessors
        return types.MethodType(getattr(self._cls, name), self)


def main():
    classes = import_reference()
    rs = np.random.RandomState(2024)
    out = {}
    # ---- get_laser_observation: 512 raw ranges incl. NaN / inf -> beam_num values (stage_world1.py:122-140)
    scan = rs.uniform(0.1, 5.9, 512)
    scan[[3, 77, 300]] = np.nan
    scan[[0, 511, 256]] = np.inf
    out['scan_raw'] = scan
    for name, cls in classes.items():
        for nb in (512, 360, 180):
            o = Bare(cls, scan=scan.copy(), beam_mum=nb)
            out[f'obs_{name}_{nb}'] = np.asarray(o.get_laser_observation(), np.float64)
    # ---- get_local_goal (stage_world1.py:155-160)
    K = 64
    state = np.stack([rs.uniform(-9, 9, K), rs.uniform(-9, 9, K), rs.uniform(-np.pi, np.pi, K)], 1)
    goals = rs.uniform(-9, 9, (K, 2))
    out['lg_state'], out['lg_goal'] = state, goals
    for name, cls in classes.items():
        out[f'local_goal_{name}'] = np.asarray(
            [Bare(cls, state_GT=list(state[i]), goal_point=list(goals[i])).get_local_goal() for i in range(K)], np.float64)
    # ---- get_reward_and_terminate (stage_world1.py:180-211, stage_world2.py:175-208, circle_world.py:171-203)
    # cases: x, y, goal_x, goal_y, previous distance, crash flag, ground-truth w, step counter t
    cases = []
    for i in range(40):
        x, y = rs.uniform(-5, 5, 2)
        gx, gy = x + rs.uniform(-4, 4), y + rs.uniform(-4, 4)
        d_prev = float(np.hypot(gx - x, gy - y) + rs.uniform(-0.1, 0.1))
        cases.append([x, y, gx, gy, d_prev, 0, rs.uniform(-1, 1), int(rs.randint(1, 100))])
    cases += [
        [0.0, 0.0, 0.3, 0.0, 0.4, 0, 0.0, 5],           # reach goal
        [0.0, 0.0, 0.3, 0.0, 0.4, 1, 0.0, 5],           # reach goal AND crashed: 15 - 15
        [1.0, 1.0, 5.0, 1.0, 4.1, 1, 0.0, 9],           # crashed
        [1.0, 1.0, 5.0, 1.0, 4.0, 0, 0.9, 9],           # |w| = 0.9: penalised only by circle_world (0.7)
        [1.0, 1.0, 5.0, 1.0, 4.0, 0, -1.0, 9],
        [1.0, 1.0, 5.0, 1.0, 4.0, 0, 0.0, 151],         # time-out of stage 1 (t > 150)
        [1.0, 1.0, 5.0, 1.0, 4.0, 0, 0.0, 150],
        [1.0, 1.0, 5.0, 1.0, 4.0, 0, 0.0, 201],         # time-out of stage 2 (t > 200)
        [1.0, 1.0, 5.0, 1.0, 4.0, 1, 0.0, 10001],       # crashed and timed out: the later test wins the result string
        [0.0, 0.0, 0.3, 0.0, 0.0, 0, 0.0, 1],           # first step after generate_goal_point of stage 2 / circle (pre_distance 0)
    ]
    cases = np.asarray(cases, np.float64)
    out['reward_cases'] = cases
    for name, cls in classes.items():
        res = []
        for c in cases:
            o = Bare(cls, scan=np.full(512, 6.0), beam_mum=512, state_GT=[c[0], c[1], 0.3], speed_GT=[0.2, c[6]],
                     goal_point=[c[2], c[3]], distance=c[4], is_crashed=int(c[5]))
            r, term, result = o.get_reward_and_terminate(int(c[7]))
            res.append([r, float(term), RESULT_CODE[result], o.distance, o.pre_distance])
        out[f'reward_{name}'] = np.asarray(res, np.float64)
    # ---- generate_goal_point bookkeeping: stage 1 sets pre_distance = distance to the new goal, stage 2 / circle set 0
    # (stage_world1.py:171-177, stage_world2.py:164-171, circle_world.py:164-167); random draws from the reference's own
    # rejection loops (np.random seeded), recorded with the pose they were drawn for
    np.random.seed(7)
    pre = {}
    o = Bare(classes['stage1'], state_GT=[1.0, -2.0, 0.5])
    pts1, pre1 = [], []
    for i in range(200):
        o.state_GT = [float(rs.uniform(-6, 6)), float(rs.uniform(-6, 6)), 0.1]
        o.generate_goal_point()
        pts1.append(o.state_GT[:2] + list(o.goal_point))
        pre1.append([o.pre_distance, o.distance])
    out['goal_stage1'] = np.asarray(pts1, np.float64)             # x, y, goal_x, goal_y
    out['goal_stage1_pre'] = np.asarray(pre1, np.float64)
    out['pose_stage1'] = np.asarray([o.generate_random_pose() for _ in range(400)], np.float64)
    o2 = Bare(classes['stage2'], index=40, state_GT=[12.0, -5.0, 0.0])
    pts2 = []
    for i in range(200):
        o2.generate_goal_point()
        pts2.append(list(o2.goal_point) + [o2.pre_distance, o2.distance])
    out['goal_stage2_random'] = np.asarray(pts2, np.float64)
    out['pose_stage2_random'] = np.asarray([o2.generate_random_pose() for _ in range(400)], np.float64)
    o2t = Bare(classes['stage2'], index=5, state_GT=[0.0, 0.0, 0.0])
    o2t.generate_goal_point()
    out['goal_stage2_table5'] = np.asarray(list(o2t.goal_point) + [o2t.pre_distance], np.float64)
    oc = Bare(classes['circle'], index=7, state_GT=[0.0, 0.0, 0.0])
    oc.generate_goal_point()
    out['goal_circle_table7'] = np.asarray(list(oc.goal_point) + [oc.pre_distance], np.float64)
    path = os.path.join(ROOT, 'tests', 'golden', 'env_golden.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays')
    for k in ('reward_stage1', 'reward_circle'):
        print(k, out[k][-10:, :3])


if __name__ == '__main__':
    main()
