"""Circle-swap env client (mirror of /root/reference/circle_world.py): 50 robots on a 25 m circle,
antipodal goals, |w| > 0.7 penalty, timeout 10000."""
from .stage_world import StageWorld as _Base


class StageWorld(_Base):
    def __init__(self, beam_num, index=0, num_env=50, **kw):
        kw.setdefault('scenario', 'circle')
        super().__init__(beam_num, index, num_env, **kw)
