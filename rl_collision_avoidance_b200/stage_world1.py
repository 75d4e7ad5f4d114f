"""Stage-1 env client (mirror of /root/reference/stage_world1.py): open rink arena, 24 robots per
world, random spawn/goal, timeout 150."""
from .stage_world import StageWorld as _Base


class StageWorld(_Base):
    def __init__(self, beam_num, index=0, num_env=24, **kw):
        kw.setdefault('scenario', 'stage1')
        super().__init__(beam_num, index, num_env, **kw)
