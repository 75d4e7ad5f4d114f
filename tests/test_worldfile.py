"""World-file + bitmap loader on a tiny synthetic world (no reference files needed)."""
import numpy as np

import pytest

from rl_collision_avoidance_b200.worldfile import WorldFileError, bitmap_rects, load_world, parse_worldfile

WORLD = '''
# comment
resolution 0.5
define wall model ( color "gray30" boundary 1 ranger_return 1 )
define laser ranger ( sensor( fov 180 range [0.0 6.0] samples 512 ) )
define bot position ( size [0.44 0.38 0.22] drive "diff" laser( pose [0 0 0 0] ) )
wall ( name "m" size [8.0 4.0 0.8] pose [0 0 0 0] bitmap "m.png" )
bot( pose [1.00 -0.50 0.00 90.00])
bot( pose [-2.0 0.0 0.0 180.0] )
define obstacle position ( ranger_return 1 )
obstacle( pose [2 1 0 0] size [1.0 1.0 0.8]
  block( points 4 point[0] [0 0] point[1] [1 0] point[2] [1 1] point[3] [0 1] z [0 1] ) )
'''


def write_world(tmp_path):
    from PIL import Image
    img = np.full((8, 16), 255, np.uint8)
    img[0, :] = 0          # top wall
    img[:, 0] = 0          # left wall
    img[4, 6:10] = 0       # a bar
    Image.fromarray(img, 'L').save(tmp_path / 'm.png')
    (tmp_path / 't.world').write_text(WORLD)
    return str(tmp_path / 't.world')


def test_parser(tmp_path):
    g, ents = parse_worldfile(write_world(tmp_path))
    assert g['resolution'] == 0.5
    kinds = [e.kind for e in ents]
    assert kinds == ['wall', 'bot', 'bot', 'obstacle']
    assert ents[0].get('bitmap') == 'm.png' and ents[0].get('boundary') == 1.0   # inherited from the define
    assert ents[1].get('pose') == [1.0, -0.5, 0.0, 90.0]
    assert ents[3].children[0].get('point[2]') == [1.0, 1.0]


def test_bitmap_rects_cover_dark_pixels():
    img = np.full((6, 6), 255, np.uint8)
    img[1:4, 2:5] = 0
    rects = bitmap_rects(img)
    assert rects == [(2.0, 6 - 1 - (1 + 3), 3.0, 3.0)]


def test_load_world(tmp_path):
    m = load_world(write_world(tmp_path), pitch_align=4)
    assert m.resolution == 0.5 and m.init_poses.shape == (2, 3)
    assert np.allclose(m.init_poses[0], [1.0, -0.5, np.pi / 2])
    occ = m.cells > 0
    # model spans [-4,4] x [-2,2] m -> cells -8..8 x -4..4 (+1 margin) ; boundary closes the box
    def cell(x, y):
        return occ[int(np.floor(y * 2)) + m.origin_cy, int(np.floor(x * 2)) + m.origin_cx]
    assert cell(-3.9, 1.9) and cell(3.9, 1.9) and cell(-3.9, -1.9) and cell(3.9, -1.9)   # corners
    assert cell(0.0, 1.9) and cell(-3.9, 0.0) and cell(3.9, 0.0) and cell(0.0, -1.95)    # four sides
    assert not cell(-2.0, 0.9) and not cell(2.9, -1.0)                                    # free interior
    assert cell(2.0, 1.0) or cell(1.6, 0.6)                                               # polygon obstacle outline
    assert m.cells.shape[1] % 4 == 0


@pytest.mark.parametrize('text,what', [
    ('resolution 0.5\nbot( pose [1 2 0 90]', 'unbalanced'),                 # missing ')'
    ('resolution 0.5\nbot( pose [1 2 0 90 )', 'unbalanced'),                # missing ']' swallows the rest
    ('define bot position size [1 1 1] )', "expected '('"),                 # define without a body
])
def test_malformed_world_files_raise_a_clear_error(tmp_path, text, what):
    f = tmp_path / 'bad.world'
    f.write_text(text)
    with pytest.raises(WorldFileError, match=what.replace('(', r'\(')):
        parse_worldfile(str(f))


def test_world_without_static_geometry_is_rejected(tmp_path):
    f = tmp_path / 'empty.world'
    f.write_text('resolution 0.2\ndefine laser ranger ( )\ndefine bot position ( laser( ) )\nbot( pose [0 0 0 0] )\n')
    with pytest.raises(WorldFileError, match='no static geometry'):
        load_world(str(f))


def test_degenerate_polygon_obstacle_is_rejected(tmp_path):
    from PIL import Image
    Image.fromarray(np.zeros((4, 4), np.uint8), 'L').save(tmp_path / 'm.png')
    f = tmp_path / 'deg.world'
    f.write_text('resolution 0.5\ndefine wall model ( )\nwall( size [2 2 1] pose [0 0 0 0] bitmap "m.png" )\n'
                 'define obstacle position ( )\n'
                 'obstacle( pose [0 0 0 0] size [1 1 1] block( points 2 point[0] [0 0] point[1] [1 0] ) )\n')
    with pytest.raises(WorldFileError, match='degenerate polygon'):
        load_world(str(f))
