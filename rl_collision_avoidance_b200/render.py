"""Trajectory / scan dump and offline renderer: the replacement for the Stage GUI and the
reference's doc/*.gif.  Draws one world of a batched StageWorld — static map, robot footprints with headings, goals,
and the lidar returns of selected robots — into a PIL image; `record()` collects frames into an animated GIF and a
.npz trajectory (poses, goals, flags per tick) for debugging parity failures."""
from __future__ import annotations

import math

import numpy as np


def _to_px(m, x, y, scale):
    cx = (x / m.resolution + m.origin_cx) * scale
    cy = (m.grid_h - (y / m.resolution + m.origin_cy)) * scale      # image row 0 = +y
    return cx, cy


def render_world(env, world=0, scale=None, scan_robots=(0,), size=720):
    """PIL.Image of world `world`.  Lidar endpoints are reconstructed from env.obs (scan/6-0.5)."""
    from PIL import Image, ImageDraw
    m = env.sc.map
    scale = scale or max(1e-3, size / max(m.grid_w, m.grid_h))
    W, H = int(m.grid_w * scale), int(m.grid_h * scale)
    occ = (np.asarray(m.cells) > 0)
    if scale >= 1:
        img = np.kron(occ[::-1], np.ones((int(scale), int(scale)), bool))
        im = Image.fromarray(np.where(img, 60, 255).astype(np.uint8), 'L').convert('RGB').resize((W, H))
    else:
        step = int(math.ceil(1.0 / scale))
        h2, w2 = occ.shape[0] // step * step, occ.shape[1] // step * step
        coarse = occ[:h2, :w2].reshape(h2 // step, step, w2 // step, step).any(axis=(1, 3))
        im = Image.fromarray(np.where(coarse[::-1], 60, 255).astype(np.uint8), 'L').convert('RGB').resize((W, H))
    d = ImageDraw.Draw(im)
    R = env.num_env
    pose = env.state['pose'][world * R:(world + 1) * R].cpu().numpy()
    goal = env.state['goal'][world * R:(world + 1) * R].cpu().numpy()
    meta = env.state['meta'][world * R:(world + 1) * R].cpu().numpy()
    obs = env.obs[world * R:(world + 1) * R].cpu().numpy()
    hl, hw = 0.22, 0.19
    for r in range(R):
        x, y, th = pose[r, :3]
        c, s = math.cos(th), math.sin(th)
        pts = [_to_px(m, x + a * c - b * s, y + a * s + b * c, scale) for a, b in ((-hl, -hw), (hl, -hw), (hl, hw), (-hl, hw))]
        col = (220, 40, 40) if meta[r, 2] else (30, 90, 220)
        d.polygon(pts, outline=col, fill=col if scale * 2 * hl / m.resolution < 4 else None)
        hx, hy = _to_px(m, x + 0.4 * c, y + 0.4 * s, scale)
        cx, cy = _to_px(m, x, y, scale)
        d.line([cx, cy, hx, hy], fill=col)
        gx, gy = _to_px(m, goal[r, 0], goal[r, 1], scale)
        d.line([gx - 3, gy, gx + 3, gy], fill=(20, 160, 60))
        d.line([gx, gy - 3, gx, gy + 3], fill=(20, 160, 60))
    nb = obs.shape[1]
    for r in scan_robots:
        if r >= R:
            continue
        x, y, th = pose[r, :3]
        rng = (obs[r] + 0.5) * 6.0
        for b in range(0, nb, max(1, nb // 128)):
            a = th - math.pi / 2 + b * math.pi / (nb - 1)
            ex, ey = _to_px(m, x + rng[b] * math.cos(a), y + rng[b] * math.sin(a), scale)
            d.point([ex, ey], fill=(240, 140, 0))
    return im


def record(env, policy_step, ticks, world=0, gif_path=None, npz_path=None, every=1, **kw):
    """Run `ticks` ticks; policy_step(env) must advance the env by one tick.  Saves an animated GIF and/or a
    trajectory dump (poses (T,R,4), goals (T,R,4), flags (T,R,4))."""
    frames, poses, goals, flags = [], [], [], []
    R = env.num_env
    for t in range(ticks):
        policy_step(env)
        poses.append(env.state['pose'][world * R:(world + 1) * R].cpu().numpy().copy())
        goals.append(env.state['goal'][world * R:(world + 1) * R].cpu().numpy().copy())
        flags.append(env.flags[world * R:(world + 1) * R].cpu().numpy().copy())
        if gif_path and t % every == 0:
            frames.append(render_world(env, world, **kw))
    if gif_path and frames:
        frames[0].save(gif_path, save_all=True, append_images=frames[1:], duration=100, loop=0)
    if npz_path:
        np.savez_compressed(npz_path, pose=np.stack(poses), goal=np.stack(goals), flags=np.stack(flags))
    return len(frames)
