"""Host-side helpers mirroring /root/reference/model/utils.py:65-97 (tables live in scenarios.py)."""
from __future__ import annotations

import bisect
import math

import numpy as np
import torch

GROUP_REFER = [0, 6, 10, 15, 19, 24, 34, 44]          # model/utils.py:83


def get_filter_index(d_list):
    """Flat indices (num_env*t + i) of the 2nd and later consecutive True in each column, with the
    reference's quirk that the run counter is NOT reset between columns (model/utils.py:65-78;
    SURVEY.md App. D.6).  Accepts a (T, N) bool tensor or array; returns a python list like the reference."""
    d = d_list.detach().cpu().numpy() if isinstance(d_list, torch.Tensor) else np.asarray(d_list)
    d = d.astype(bool)
    step, num_env = d.shape
    col_major = d.T.reshape(-1)                        # the reference's loop order: env outer, step inner
    # run length of consecutive True ending at each position (counter carried across columns)
    idx = np.arange(col_major.size)
    last_false = np.maximum.accumulate(np.where(~col_major, idx, -1))
    run = idx - last_false
    hit = col_major & (run >= 2)
    pos = np.nonzero(hit)[0]
    i, j = pos // step, pos % step
    return list((num_env * j + i).astype(np.int64))


def get_group_terminal(terminal_list, index=None, refer=GROUP_REFER):
    """Group barrier of stage 2 (model/utils.py:81-87, ppo_stage2.py:105-106).  `terminal_list` is (N,) or
    (num_worlds, N) bool; returns for every agent whether ALL members of its group have terminated
    (the reference evaluates this per robot `index`; pass index to get that single bool)."""
    t = terminal_list if isinstance(terminal_list, torch.Tensor) else torch.as_tensor(np.asarray(terminal_list))
    t2 = t.reshape(-1, refer[-1]).bool()
    out = torch.empty_like(t2)
    for a, b in zip(refer[:-1], refer[1:]):
        out[:, a:b] = t2[:, a:b].all(dim=1, keepdim=True)
    out = out.reshape(t.shape)
    if index is not None:
        r = bisect.bisect(refer, index)
        return bool(out.reshape(-1)[index]) if r > 0 else False
    return out


def log_normal_density(x, mean, log_std, std):
    """returns gaussian density given x on log scale (model/utils.py:90-97)"""
    variance = std.pow(2)
    log_density = -(x - mean).pow(2) / (2 * variance) - 0.5 * math.log(2 * math.pi) - log_std
    return log_density.sum(dim=-1, keepdim=True)
