"""Build the in-repo scenario assets from the reference's world files and tables.

Run HERE (where /root/reference exists):  python tools/build_assets.py
Outputs (committed; they travel to the GPU box, /root/reference does not):
  rl_collision_avoidance_b200/assets/{stage1,stage2,circle}_map.npz  static occupancy grids + agent start poses
  rl_collision_avoidance_b200/assets/scenarios.json                  spawn/goal tables (model/utils.py:6-63)
The maps are produced by OUR loader (rl_collision_avoidance_b200/worldfile.py) from
/root/reference/worlds/*.world + *.png; the tables by importing the reference's
model/utils.py read-only and calling its table functions.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('RLCA_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)

from rl_collision_avoidance_b200.worldfile import load_world, save_map  # noqa: E402


def main():
    out = os.path.join(ROOT, 'rl_collision_avoidance_b200', 'assets')
    os.makedirs(out, exist_ok=True)
    for name in ('stage1', 'stage2', 'circle'):
        m = load_world(os.path.join(REF, 'worlds', name + '.world'))
        save_map(m, os.path.join(out, name + '_map.npz'))
        print(name, m.cells.shape, 'occupied', int((m.cells > 0).sum()), 'agents', len(m.init_poses))
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir('/tmp')
    from model import utils as ref_utils  # the reference's own tables
    os.chdir(cwd)
    tabs = {
        'stage2': {
            'init_pose': [list(map(float, ref_utils.get_init_pose(i))) for i in range(44)],
            'goal_point': [list(map(float, ref_utils.get_goal_point(i))) for i in range(34)],
            'groups': [0, 6, 10, 15, 19, 24, 34, 44],   # model/utils.py:83
        },
        'circle': {
            'init_pose': [list(map(float, ref_utils.test_init_pose(i))) for i in range(50)],
            'goal_point': [list(map(float, ref_utils.test_goal_point(i))) for i in range(50)],
        },
    }
    with open(os.path.join(out, 'scenarios.json'), 'w') as f:
        json.dump(tabs, f, indent=1)
    print('wrote scenarios.json')


if __name__ == '__main__':
    main()
