"""Offline renderer (the replacement for the Stage GUI): a known 3-robot frame, no GPU."""
import numpy as np
import torch


class _Map:
    resolution = 0.1
    grid_w, grid_h = 100, 80
    origin_cx, origin_cy = 50, 40
    cells = np.zeros((80, 100), np.uint8)


class _Scenario:
    map = _Map()


class _Env:
    """Duck-typed stand-in for StageWorld: what render_world reads."""
    sc = _Scenario()
    num_env = 3

    def __init__(self):
        self.state = {
            'pose': torch.tensor([[-3.0, 0.0, 0.0, 0.0], [0.0, 2.0, np.pi / 2, 0.0], [3.0, -2.0, np.pi, 0.0]]),
            'goal': torch.tensor([[3.0, 0.0, 0.0, 0.0], [0.0, -2.0, 0.0, 0.0], [-3.0, 2.0, 0.0, 0.0]]),
            'meta': torch.tensor([[1, 0, 0, 0], [1, 0, 1, 0], [1, 0, 0, 0]], dtype=torch.int32),      # robot 1 crashed
        }
        self.obs = torch.full((3, 512), 0.5)                   # every beam at max range
        self.obs[0, 250:262] = (1.0 / 6.0) - 0.5               # robot 0 sees something 1 m straight ahead
        self.flags = torch.zeros(3, 4, dtype=torch.uint8)


def _px(x, y, scale=4.0):
    m = _Map
    return int((x / m.resolution + m.origin_cx) * scale), int((m.grid_h - (y / m.resolution + m.origin_cy)) * scale)


def test_render_known_three_robot_frame(tmp_path):
    from rl_collision_avoidance_b200.render import record, render_world
    _Map.cells[:2, :] = 1                                       # a wall along the bottom (y = -4 m)
    env = _Env()
    im = render_world(env, world=0, scale=4.0, scan_robots=(0,))
    assert im.size == (400, 320)
    a = np.asarray(im).astype(int)
    blue = (a[..., 2] > 180) & (a[..., 0] < 80)
    red = (a[..., 0] > 180) & (a[..., 2] < 80)
    green = (a[..., 1] > 130) & (a[..., 0] < 60)
    orange = (a[..., 0] > 200) & (a[..., 1] > 100) & (a[..., 1] < 180) & (a[..., 2] < 40)
    dark = (a.max(-1) < 90)

    def near(mask, x, y, r=8):
        cx, cy = _px(x, y)
        return mask[max(0, cy - r):cy + r, max(0, cx - r):cx + r].any()
    assert near(blue, -3.0, 0.0) and near(blue, 3.0, -2.0)      # robots 0 and 2: normal colour at their poses
    assert near(red, 0.0, 2.0) and not near(blue, 0.0, 2.0, r=3)   # robot 1: crashed colour
    assert near(green, 3.0, 0.0) and near(green, 0.0, -2.0) and near(green, -3.0, 2.0)      # the three goal crosses
    assert near(orange, -2.0, 0.0, r=6)                          # robot 0's lidar return 1 m ahead (heading 0)
    assert dark[-6:, :].mean() > 0.9 and dark[:100, :].mean() < 0.05      # the wall is at the bottom of the image (y flipped)
    # record(): frames + trajectory dump
    n = record(env, lambda e: None, ticks=3, world=0, gif_path=str(tmp_path / 'r.gif'), npz_path=str(tmp_path / 'r.npz'),
               every=1, scale=4.0)
    assert n == 3 and (tmp_path / 'r.gif').stat().st_size > 0
    d = np.load(tmp_path / 'r.npz')
    assert d['pose'].shape == (3, 3, 4) and d['goal'].shape == (3, 3, 4)
