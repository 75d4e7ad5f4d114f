"""CPU property tests of the integer-walk identities the CUDA kernels rely on.

1. Closed-form position inside a Cohen walk: with a = 2ax, b = 2ay, D = a + b and the (negated) error term N0 at the
   current cell, the number of x-steps after k further steps is i(k) = max(0, ceil((N0 + a (k-1)) / D)).  The
   global-grid path (circle.world) uses it to jump over empty 16 x 16 tiles; the jump must land exactly where the
   cell-by-cell walk would be, with the same error term.
2. A walk of n = ax + ay steps ends exactly on start + (idx, idy) and visits every cell once, so the march loop may test
   for the end cell once per PAIR of steps after taking a single step first when n is odd.
3. Beams that truncate to the same end point (idx, idy) visit the same cells (the dedupe of phase 1), whatever their
   exact direction.
"""
import math

import numpy as np

T = 16       # tile edge of the coarse bitmap (TILE_SHIFT = 4 in csrc/rlca_env.cu)


def walk_cells(cx, cy, idx, idy):
    """Reference: the cell-by-cell walk; yields (x, y, nexy) BEFORE each of the n steps."""
    sx, sy = (idx > 0) - (idx < 0), (idy > 0) - (idy < 0)
    ax, ay = abs(idx), abs(idy)
    nexy = ax - ay                       # negated error term: x-step iff nexy > 0
    out = []
    for _ in range(ax + ay):
        out.append((cx, cy, nexy))
        if nexy > 0:
            cx += sx
            nexy -= 2 * ay
        else:
            cy += sy
            nexy += 2 * ax
    return out, (cx, cy, nexy)


def tile_exit_jump(cx, cy, nexy, n, sx, sy, ax, ay):
    """The kernel's jump out of an empty tile (march_walk<true>), restated: returns the state after the jump."""
    a, b = 2 * ax, 2 * ay
    D = a + b
    dxb = T - (cx & (T - 1)) if sx > 0 else (cx & (T - 1)) + 1
    dyb = T - (cy & (T - 1)) if sy > 0 else (cy & (T - 1)) + 1
    k = n
    if a > 0:
        Rx = D * (dxb - 1) - nexy
        k = min(k, 1 if Rx < 0 else Rx // a + 2)
    if b > 0:
        Ry = D * dyb - a + nexy
        k = min(k, 1 if Ry <= b else (Ry + b - 1) // b)
    num = nexy + a * (k - 1)
    i = (num + D - 1) // D if num > 0 else 0
    j = k - i
    return cx + sx * i, cy + sy * j, nexy + a * j - b * i, n - k, k


def test_closed_form_x_steps_and_tile_jumps_match_the_stepwise_walk():
    rng = np.random.default_rng(0)
    checked_jumps = 0
    for _ in range(3000):
        R = int(rng.integers(1, 700))
        ang = rng.uniform(-math.pi, math.pi)
        idx, idy = int(R * math.cos(ang)), int(R * math.sin(ang))
        if idx == 0 and idy == 0:
            continue
        cx0, cy0 = int(rng.integers(1000, 2000)), int(rng.integers(1000, 2000))
        cells, end = walk_cells(cx0, cy0, idx, idy)
        assert end[:2] == (cx0 + idx, cy0 + idy)                       # identity 2: the walk ends on the end cell
        assert len(set(c[:2] for c in cells)) == len(cells)            # ... and never visits a cell twice
        sx, sy = (idx > 0) - (idx < 0), (idy > 0) - (idy < 0)
        ax, ay = abs(idx), abs(idy)
        a, b, D = 2 * ax, 2 * ay, 2 * ax + 2 * ay
        # identity 1 at the start cell: i(k) for every k
        N0 = ax - ay
        xs = 0
        for k in range(1, len(cells) + 1):
            x, y, _ = cells[k] if k < len(cells) else end
            num = N0 + a * (k - 1)
            i = (num + D - 1) // D if num > 0 else 0
            assert abs(x - cx0) == i and abs(y - cy0) == k - i
        # tile jumps from random positions along the walk land on the stepwise state
        for s in rng.integers(0, len(cells), size=4):
            cx, cy, nexy = cells[s]
            n = len(cells) - s
            nx, ny, nn, nleft, k = tile_exit_jump(cx, cy, nexy, n, sx, sy, ax, ay)
            assert k >= 1 and nleft == n - k
            ref = cells[s + k] if s + k < len(cells) else end
            assert (nx, ny, nn) == ref
            # every cell skipped by the jump lies in the tile the jump started in (so an empty tile hides nothing) ...
            for q in range(s, s + k):
                assert (cells[q][0] >> 4, cells[q][1] >> 4) == (cx >> 4, cy >> 4)
            # ... and the jump is maximal: the landing cell is outside that tile unless the walk ended
            if s + k < len(cells):
                assert (nx >> 4, ny >> 4) != (cx >> 4, cy >> 4)
            checked_jumps += 1
    assert checked_jumps > 5000


def test_pair_stepping_meets_the_end_cell_exactly():
    """The march loop's structure: odd n -> one single step, then pairs with ONE end test per pair."""
    rng = np.random.default_rng(1)
    for _ in range(2000):
        idx, idy = int(rng.integers(-40, 41)), int(rng.integers(-40, 41))
        if idx == 0 and idy == 0:
            continue
        cells, end = walk_cells(0, 0, idx, idy)
        n = len(cells)
        pos = 0
        visited = []
        if n & 1:
            visited.append(cells[pos][:2])
            pos += 1
        while pos < n:                       # the pair loop: test, step, test, step, end check
            visited.append(cells[pos][:2])
            visited.append(cells[pos + 1][:2])
            pos += 2
        assert pos == n and visited == [c[:2] for c in cells]


def test_beams_with_the_same_truncated_end_point_share_the_walk():
    rng = np.random.default_rng(2)
    R = 30.0                                 # range_cells of stage 1/2: 6 m at 0.2 m cells
    groups = {}
    for ang in rng.uniform(-math.pi, math.pi, 20000):
        key = (int(R * math.cos(ang)), int(R * math.sin(ang)))
        cells, _ = walk_cells(50, 50, *key)
        sig = tuple(c[:2] for c in cells)
        assert groups.setdefault(key, sig) == sig
    # about five of 512 adjacent beams share a key at this resolution
    fov = math.pi
    keys = [(int(R * math.cos(-fov / 2 + i * fov / 511)), int(R * math.sin(-fov / 2 + i * fov / 511))) for i in range(512)]
    distinct = 1 + sum(keys[i] != keys[i - 1] for i in range(1, 512))
    assert 80 <= distinct <= 130

