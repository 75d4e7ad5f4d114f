"""Stage-2 env client (mirror of /root/reference/stage_world2.py): testenv map + polygon obstacles,
44 robots per world, table spawn/goal for robots 0..33, random region for 34..43, timeout 200."""
from .stage_world import StageWorld as _Base


class StageWorld(_Base):
    def __init__(self, beam_num, index=0, num_env=44, **kw):
        kw.setdefault('scenario', 'stage2')
        super().__init__(beam_num, index, num_env, **kw)
