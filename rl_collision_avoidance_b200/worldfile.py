"""Stage world-file + bitmap loader -> static occupancy grid and robot start table.

Input side of the hot path.  Reads the subset of the
Stage worldfile grammar the reference's scenarios use
(/root/reference/worlds/stage1.world:3-130, stage2.world:3-297,
circle.world:3-155): ``resolution``, ``define <name> <base> (...)`` macros,
``floorplan(size, pose, bitmap, boundary)``, ``agent(pose [...])`` and
``obstacle(pose, size, block(points ...))``.

Rasterisation follows libstage as recalled in SURVEY.md App. A.3/A.4/A.5
(source absent -> the grid produced here is *the* map; the simulator kernels
and the oracle both take it as an input):

* bitmap -> rectangles: a pixel is an obstacle when its first channel <= 127;
  greedy run/column merging, y inverted (image row 0 = +y);
* the rectangles' bounding box is scaled to the model ``size`` and centred on
  the model ``pose``; ``boundary 1`` adds four thin rectangles round that box;
* only polygon EDGES are rasterised, with the Cohen integer line between
  ``floor(x * ppm)`` cells (start cell included, end cell excluded);
  deviation d1: an edge whose two ends share a cell still marks that cell, so
  sub-cell features are not lost.

Cell value 254 = static obstacle (the simulator's owner-grid convention).
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field

import numpy as np

CELL_STATIC = 254


# --------------------------------------------------------------------------- parser
_TOKEN = re.compile(r'"[^"]*"|[A-Za-z_]\w*\[\d+\]|\(|\)|\[|\]|[^\s()\[\]"]+')


def _tokenize(text: str):
    toks = []
    for line in text.splitlines():
        line = line.split('#', 1)[0]
        toks.extend(_TOKEN.findall(line))
    return toks


@dataclass
class Entity:
    kind: str
    props: dict = field(default_factory=dict)
    children: list = field(default_factory=list)

    def get(self, key, default=None):
        return self.props.get(key, default)


def _parse_value(toks, i):
    """Parse one property value starting at toks[i]; returns (value, next_i)."""
    t = toks[i]
    if t == '[':
        vals = []
        i += 1
        while toks[i] != ']':
            vals.append(_scalar(toks[i]))
            i += 1
        return vals, i + 1
    return _scalar(t), i + 1


def _scalar(t):
    if t.startswith('"'):
        return t[1:-1]
    try:
        return float(t)
    except ValueError:
        return t


def _parse_body(toks, i, ent: Entity, macros):
    """Parse `( ... )` body of an entity; toks[i] is the token after '('."""
    while toks[i] != ')':
        name = toks[i]
        if toks[i + 1] == '(':
            child = _instantiate(name, macros)
            i = _parse_body(toks, i + 2, child, macros)
            ent.children.append(child)
        else:
            val, i = _parse_value(toks, i + 1)
            ent.props[name] = val
    return i + 1


def _instantiate(name, macros) -> Entity:
    if name in macros:
        base = macros[name]
        ent = Entity(kind=name, props=dict(base.props), children=list(base.children))
        ent.base_kind = getattr(base, 'base_kind', base.kind)
    else:
        ent = Entity(kind=name)
        ent.base_kind = name
    return ent


class WorldFileError(ValueError):
    """Malformed world file (unbalanced brackets, bad define, geometry that cannot be rasterised)."""


def parse_worldfile(path: str):
    """Returns (globals dict, list of top-level entities).  Raises WorldFileError on malformed input."""
    with open(path, 'r') as f:
        toks = _tokenize(f.read())
    try:
        return _parse_tokens(toks)
    except IndexError:
        raise WorldFileError(f"{path}: unexpected end of file (unbalanced '(' or '[')") from None


def _parse_tokens(toks):
    macros = {}
    globals_ = {}
    entities = []
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == 'define':
            name, base = toks[i + 1], toks[i + 2]
            if toks[i + 3] != '(':
                raise WorldFileError(f"define {name}: expected '(' after the base model, got {toks[i + 3]!r}")
            ent = _instantiate(base, macros)
            ent.kind = name
            i = _parse_body(toks, i + 4, ent, macros)
            macros[name] = ent
        elif i + 1 < len(toks) and toks[i + 1] == '(':
            ent = _instantiate(t, macros)
            i = _parse_body(toks, i + 2, ent, macros)
            entities.append(ent)
        else:
            val, i = _parse_value(toks, i + 1)
            globals_[t] = val
    return globals_, entities


# --------------------------------------------------------------------------- raster
def bitmap_rects(img: np.ndarray, threshold: int = 127):
    """Greedy dark-pixel -> rectangle merge (rotrects_from_image_file, App. A.4).

    img: (H, W) first-channel uint8.  Returns list of (x, y, w, h) in pixel
    units with y already inverted (conventional, +y up)."""
    free = img > threshold            # 'set' (white) pixels
    free = free.copy()
    H, W = free.shape
    # depth[y, x] = number of consecutive dark pixels downward from (x, y), with the
    # reference quirk that the scan stops at row H-1 (the bottom row never counts).
    rects = []
    for y in range(H):
        x = 0
        row_free = free[y]
        while x < W:
            if row_free[x]:
                x += 1
                continue
            startx = x
            rheight = H
            while x < W and not free[y, x]:
                yy = y
                while (not free[yy, x]) and yy < H - 1:
                    yy += 1
                if yy - y < rheight:
                    rheight = yy - y
                x += 1
            free[y:y + rheight, startx:x] = True
            if rheight == 0:
                # zero-height run on the last row: consumed nothing; skip past it
                free[y, startx:x] = True
            rects.append((float(startx), float(H - 1 - (y + rheight)), float(x - startx), float(rheight)))
    return rects


def _line_cells(x0, y0, x1, y1):
    dx, dy = x1 - x0, y1 - y0
    sx = (dx > 0) - (dx < 0)
    sy = (dy > 0) - (dy < 0)
    ax, ay = abs(dx), abs(dy)
    bx, by = 2 * ax, 2 * ay
    exy = ay - ax
    n = ax + ay
    gx, gy = x0, y0
    out = []
    if n == 0:
        out.append((gx, gy))          # deviation d1
    while n > 0:
        out.append((gx, gy))
        if exy < 0:
            gx += sx
            exy += by
        else:
            gy += sy
            exy -= bx
        n -= 1
    return out


def _polygon_cells(pts_world, ppm):
    cells = []
    k = len(pts_world)
    ij = [(int(math.floor(p[0] * ppm)), int(math.floor(p[1] * ppm))) for p in pts_world]
    for a in range(k):
        b = (a + 1) % k
        cells.extend(_line_cells(ij[a][0], ij[a][1], ij[b][0], ij[b][1]))
    return cells


def _transform(pts, pose):
    x0, y0, th = pose
    c, s = math.cos(th), math.sin(th)
    return [(x0 + px * c - py * s, y0 + px * s + py * c) for px, py in pts]


@dataclass
class WorldMap:
    cells: np.ndarray          # (grid_h, grid_w) uint8, 254 = static obstacle
    resolution: float
    origin_cx: int             # cell column of world x = 0
    origin_cy: int
    init_poses: np.ndarray     # (R, 3) x, y, theta[rad] of the `agent` models, file order
    name: str = ''

    @property
    def grid_w(self):
        return int(self.cells.shape[1])

    @property
    def grid_h(self):
        return int(self.cells.shape[0])


def _pose_of(ent: Entity):
    p = ent.get('pose', [0.0, 0.0, 0.0, 0.0])
    return (float(p[0]), float(p[1]), math.radians(float(p[3]) if len(p) > 3 else 0.0))


def load_world(path: str, pitch_align: int = 16, margin: int = 1) -> WorldMap:
    from PIL import Image

    globals_, entities = parse_worldfile(path)
    resolution = float(globals_.get('resolution', 0.02))
    ppm = 1.0 / resolution
    marked = set()
    agents = []
    for ent in entities:
        base = getattr(ent, 'base_kind', ent.kind)
        if ent.kind == 'window':
            continue
        if ent.kind == 'floorplan' or (base == 'model' and ent.get('bitmap') is not None):
            size = ent.get('size')
            pose = _pose_of(ent)
            bmp = os.path.join(os.path.dirname(path), ent.get('bitmap'))
            img = np.asarray(Image.open(bmp))
            if img.ndim == 3:
                img = img[:, :, 0]
            rects = bitmap_rects(img)
            if not rects or size is None:
                raise WorldFileError(f'{path}: bitmap model {ent.get("name", ent.kind)!r} needs dark pixels and a size')
            xs0 = min(r[0] for r in rects)
            ys0 = min(r[1] for r in rects)
            xs1 = max(r[0] + r[2] for r in rects)
            ys1 = max(r[1] + r[3] for r in rects)
            if int(float(ent.get('boundary', 0))):
                eps = 0.01
                bw, bh = xs1 - xs0, ys1 - ys0
                rects = rects + [(xs0, ys0, eps, bh), (xs0, ys0, bw, eps),
                                 (xs0, ys1 - eps, bw, eps), (xs1 - eps, ys0, eps, bh)]
            if xs1 <= xs0 or ys1 <= ys0:
                raise WorldFileError(f'{path}: bitmap {ent.get("bitmap")!r} has a degenerate obstacle extent')
            scx = float(size[0]) / (xs1 - xs0)
            scy = float(size[1]) / (ys1 - ys0)
            offx, offy = 0.5 * (xs0 + xs1), 0.5 * (ys0 + ys1)
            for (x, y, w, h) in rects:
                loc = [((x - offx) * scx, (y - offy) * scy), ((x + w - offx) * scx, (y - offy) * scy),
                       ((x + w - offx) * scx, (y + h - offy) * scy), ((x - offx) * scx, (y + h - offy) * scy)]
                marked.update(_polygon_cells(_transform(loc, pose), ppm))
        elif base == 'position' and any(getattr(c, 'base_kind', c.kind) == 'ranger' for c in ent.children):
            # a position model carrying a ranger is a robot (worlds/stage1.world:80-105); the others are
            # never commanded and therefore static (stage2.world:105-111)
            agents.append(_pose_of(ent))
        elif base == 'position':
            # static polygon obstacle (stage2.world:169-297): blocks normalised to `size`
            size = ent.get('size', [1.0, 1.0, 1.0])
            pose = _pose_of(ent)
            blocks = [c for c in ent.children if c.kind == 'block']
            polys = []
            for b in blocks:
                npts = int(b.get('points'))
                pts = []
                for k in range(npts):
                    v = b.get(f'point[{k}]')
                    if v is not None:
                        pts.append((float(v[0]), float(v[1])))
                polys.append(pts)
            allp = [p for poly in polys for p in poly]
            if not allp:
                continue
            minx, maxx = min(p[0] for p in allp), max(p[0] for p in allp)
            miny, maxy = min(p[1] for p in allp), max(p[1] for p in allp)
            if maxx <= minx or maxy <= miny:
                raise WorldFileError(f'{path}: obstacle at {ent.get("pose")} has a degenerate polygon')
            scx = float(size[0]) / (maxx - minx)
            scy = float(size[1]) / (maxy - miny)
            offx, offy = 0.5 * (minx + maxx), 0.5 * (miny + maxy)
            for poly in polys:
                loc = [((px - offx) * scx, (py - offy) * scy) for px, py in poly]
                marked.update(_polygon_cells(_transform(loc, pose), ppm))
    if not marked:
        raise WorldFileError(f'{path}: no static geometry (floorplan bitmap or polygon obstacle) found')
    xs = [c[0] for c in marked]
    ys = [c[1] for c in marked]
    minx, maxx, miny, maxy = min(xs), max(xs), min(ys), max(ys)
    origin_cx = -minx + margin
    origin_cy = -miny + margin
    w = maxx - minx + 1 + 2 * margin
    h = maxy - miny + 1 + 2 * margin
    pitch = (w + pitch_align - 1) // pitch_align * pitch_align
    cells = np.zeros((h, pitch), dtype=np.uint8)
    idx = np.array(list(marked), dtype=np.int64)
    cells[idx[:, 1] + origin_cy, idx[:, 0] + origin_cx] = CELL_STATIC
    init = np.array(agents, dtype=np.float64).reshape(-1, 3)
    return WorldMap(cells=cells, resolution=resolution, origin_cx=int(origin_cx), origin_cy=int(origin_cy),
                    init_poses=init, name=os.path.basename(path))


def save_map(m: WorldMap, path: str):
    np.savez_compressed(path, cells=m.cells, resolution=np.float64(m.resolution),
                        origin=np.array([m.origin_cx, m.origin_cy], dtype=np.int64),
                        init_poses=m.init_poses, name=np.array(m.name))


def load_map(path: str) -> WorldMap:
    z = np.load(path, allow_pickle=False)
    return WorldMap(cells=z['cells'], resolution=float(z['resolution']), origin_cx=int(z['origin'][0]),
                    origin_cy=int(z['origin'][1]), init_poses=z['init_poses'], name=str(z['name']))
