"""Host logic of the table-driven lidar (no GPU): the walk tables the library builds (rlca_walk_tables_host, the same
code rlca_env_set_map runs) reproduce the cell-by-cell walk exactly.

Claim under test (csrc/rlca_env.cu, "Walk tables"): for a walk with truncated end point (idx, idy) from start cell c,
    first blocked cell's dominant-axis distance = min( first static hit , min over other robots' outline cells q on
                                                      the walk of dom(q - c) )
where the second term comes from the inverse lists inv[q - c] = {(slot, dom)}.  The reference semantics is the
plain integer-line walk (World::Raytrace restated; oracle/sim_oracle.c marches it)."""
import ctypes as C

import numpy as np
import pytest


def _tables(range_cells):
    from rl_collision_avoidance_b200 import _lib
    lib = _lib.load()
    kr, ns, ne = C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(lib.rlca_walk_tables_host(C.c_float(range_cells), C.byref(kr), C.byref(ns), C.byref(ne), None, None, None, None))
    kdim = 2 * kr.value + 1
    keys = np.zeros((ns.value, 2), np.int16)
    keyslot = np.zeros(kdim * kdim, np.uint16)
    off = np.zeros(kdim * kdim + 1, np.uint32)
    ent = np.zeros(max(ne.value, 1), np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.rlca_walk_tables_host(C.c_float(range_cells), C.byref(kr), C.byref(ns), C.byref(ne), p(keys), p(keyslot),
                                         p(off), p(ent)))
    return kr.value, keys, keyslot.reshape(kdim, kdim), off, ent[:ne.value]


def _walk(idx, idy):
    """Cells (relative to the start) the integer-line walk tests, in order; the end cell is not tested."""
    sx, sy = np.sign(idx), np.sign(idy)
    ax, ay = abs(idx), abs(idy)
    nexy, gx, gy = ax - ay, 0, 0
    for _ in range(ax + ay):
        yield gx, gy
        if nexy > 0:
            gx += sx
            nexy -= 2 * ay
        else:
            gy += sy
            nexy += 2 * ax


@pytest.mark.parametrize('rc', [30.0, 60.0, 7.5, 1.0])
def test_every_reachable_end_point_has_a_slot(built, rc):
    """fp32 ray directions of any heading and beam truncate to an end point that owns a slot; slots are unique."""
    kr, keys, keyslot, off, ent = _tables(rc)
    rng = np.random.default_rng(0)
    th = rng.uniform(-np.pi, np.pi, 400000).astype(np.float32)
    b = rng.uniform(-np.pi / 2, np.pi / 2, 400000)
    cb, sb = np.cos(b).astype(np.float32), np.sin(b).astype(np.float32)
    ct, st = np.cos(th).astype(np.float32), np.sin(th).astype(np.float32)
    ca = (ct * cb - st * sb).astype(np.float32)
    sa = (st * cb + ct * sb).astype(np.float32)
    idx = (np.float32(rc) * ca).astype(np.int32)               # C truncation
    idy = (np.float32(rc) * sa).astype(np.int32)
    # exact axis directions too
    idx = np.concatenate([idx, [int(rc), -int(rc), 0, 0]])
    idy = np.concatenate([idy, [0, 0, int(rc), -int(rc)]])
    assert np.abs(idx).max() <= kr and np.abs(idy).max() <= kr
    slots = keyslot[idy + kr, idx + kr]
    assert np.all(slots != 0xffff), np.stack([idx, idy], 1)[slots == 0xffff][:5]
    assert len(set(map(tuple, keys.tolist()))) == len(keys) < 0xffff
    assert len(keys) <= 8 * (int(np.ceil(rc)) + 2) + 8           # ~ the perimeter of the circle in cells
    # slots are ordered by angle: neighbouring beams read neighbouring table bytes
    ang = np.arctan2(keys[:, 1].astype(float) + 0.5 * np.sign(keys[:, 1]), keys[:, 0].astype(float) + 0.5 * np.sign(keys[:, 0]))
    assert np.all(np.diff(ang) >= -1e-12)


def test_inverse_lists_are_the_walks(built):
    kr, keys, keyslot, off, ent = _tables(30.0)
    kdim = 2 * kr + 1
    want = {}
    for s, (idx, idy) in enumerate(keys.tolist()):
        xdom = abs(idx) > abs(idy)
        for gx, gy in _walk(idx, idy):
            want.setdefault((gx, gy), set()).add((s, abs(gx) if xdom else abs(gy)))
    got = {}
    for rel in range(kdim * kdim):
        for e in ent[off[rel]:off[rel + 1]].tolist():
            got.setdefault((rel % kdim - kr, rel // kdim - kr), set()).add((e & 0xffff, e >> 16))
    assert got == want
    assert len(want[(0, 0)]) == len(keys)                          # every walk tests its own start cell first


@pytest.mark.parametrize('seed', [0, 1])
def test_tables_reproduce_the_marched_walk(built, seed):
    """Random world: static walls + blobs, 12 robots with overlapping outlines; for every robot and EVERY slot the table
    result equals the march on the owner grid (the semantics of march_walk / the oracle's scan_robot)."""
    kr, keys, keyslot, off, ent = _tables(30.0)
    kdim = 2 * kr + 1
    rng = np.random.default_rng(seed)
    W = H = 64
    static = np.zeros((H, W), bool)
    static[0, :] = static[-1, :] = static[:, 0] = static[:, -1] = True
    if seed == 1:
        static[0, 10:30] = False                                  # a gap in the wall: walks leave the map there
    for _ in range(6):
        x, y = rng.integers(5, 58, 2)
        static[y:y + rng.integers(1, 4), x:x + rng.integers(1, 4)] = True
    # padded template as the library builds it: ring of OOB (253) round the map
    OOB, ST = 253, 254
    T = np.full((H + 2, W + 2), OOB, np.uint8)
    T[1:-1, 1:-1] = np.where(static, ST, 0)
    # robots: outline cells of a 3 x 2-cell box at random places (padded coordinates), some overlapping
    robots = []
    for r in range(12):
        cx, cy = rng.integers(3, W - 3, 2) + 1
        if r % 4 == 1:
            cx, cy = robots[-1][0][0] + 1, robots[-1][0][1]          # overlaps its neighbour
        cells = {(cx + dx, cy + dy) for dx in (-1, 0, 1) for dy in (-1, 0)} - ({(cx, cy)} if r % 3 == 0 else set())
        robots.append(((cx, cy), cells))
    # owner grid (mark_outlines): 0 empty, r+1 single owner, 255 several; static / OOB cells never change
    G = T.copy()
    for r, (_, cells) in enumerate(robots):
        for (x, y) in cells:
            v = G[y, x]
            if v == 0:
                G[y, x] = r + 1
            elif v < OOB and v != r + 1:
                G[y, x] = 255
    for r, ((cx, cy), _) in enumerate(robots):
        me = r + 1
        # table side: scatter the OTHER robots' cells through the inverse lists
        hit = np.full(len(keys), 1 << 30, np.int64)
        for b, (_, cells) in enumerate(robots):
            if b == r:
                continue
            for (qx, qy) in cells:
                rx, ry = qx - cx + kr, qy - cy + kr
                if 0 <= rx <= 2 * kr and 0 <= ry <= 2 * kr and T[qy, qx] == 0:
                    rel = ry * kdim + rx
                    for e in ent[off[rel]:off[rel + 1]].tolist():
                        hit[e & 0xffff] = min(hit[e & 0xffff], e >> 16)
        for s, (idx, idy) in enumerate(keys.tolist()):
            xdom = abs(idx) > abs(idy)
            # reference: march the owner grid (start inside: the OOB ring ends the walk with a miss)
            ref = None
            for gx, gy in _walk(idx, idy):
                v = G[cy + gy, cx + gx]
                if v != 0 and v != me:
                    ref = None if v == OOB else (abs(gx) if xdom else abs(gy))
                    break
            # first static hit (what build_first_hit_kernel stores)
            fh = None
            for gx, gy in _walk(idx, idy):
                v = T[cy + gy, cx + gx]
                if v == ST:
                    fh = abs(gx) if xdom else abs(gy)
                    break
                if v == OOB:
                    break
            got = min(fh if fh is not None else 1 << 30, hit[s])
            got = None if got == 1 << 30 else got
            assert got == ref, (r, s, idx, idy, got, ref)
