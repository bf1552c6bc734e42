"""Host-side text geometry (render/text.rs, tex.rs Atlas): known answers worked out by hand from the reference's code."""
import numpy as np

import retrofire_b200 as rf
from retrofire_b200 import scenes, text


def _font():
    tex = np.arange(64 * 48 * 3, dtype=np.uint32).reshape(48, 64, 3).astype(np.uint8)
    return text.Atlas((16, 24), rf.Texture(tex))       # 4 glyphs per row, 2 rows


def test_atlas_rect_and_coords():
    """tex.rs:149-158, 178-185: glyph 5 of a 64x48 atlas with 16x24 cells is column 1 of row 1."""
    a = _font()
    assert a.rect(5) == ((16, 24), (32, 48))
    np.testing.assert_array_equal(a.coords(5), np.array([[0.25, 0.5], [0.5, 0.5], [0.25, 1.0], [0.5, 1.0]], np.float32))
    assert a.get(5).shape == (24, 16, 3) and (a.get(5) == a.texture.data[24:48, 16:32]).all()
    # no bounds check in coords (tex.rs:175-176): index 8 is below the texture
    assert a.coords(8)[2, 1] == np.float32(1.5)


def test_text_write_builds_one_quad_per_glyph():
    """text.rs:64-94, 142-162: quads advance the cursor by the glyph width; newline returns it to x = 0, y += glyph height;
    faces are (l, l+1, l+3), (l, l+3, l+2)."""
    t = text.Text(_font()).write("\x01\x02\n\x05")
    faces, verts = t.geom
    assert faces.tolist() == [[0, 1, 3], [0, 3, 2], [4, 5, 7], [4, 7, 6], [8, 9, 11], [8, 11, 10]]
    assert verts[:4, :3].tolist() == [[0, 0, 0], [16, 0, 0], [0, 24, 0], [16, 24, 0]]
    assert verts[4:8, :3].tolist() == [[16, 0, 0], [32, 0, 0], [16, 24, 0], [32, 24, 0]]
    assert verts[8:12, :3].tolist() == [[0, 24, 0], [16, 24, 0], [0, 48, 0], [16, 48, 0]]
    np.testing.assert_array_equal(verts[8:12, 3:], t.font.coords(5))
    same = text.Text(_font()).write(b"\x01\x02\n\x05").geom      # io::Write path: one glyph per byte
    assert (same[0] == faces).all() and (same[1] == verts).all()
    t.clear()
    assert t.geom[0].shape == (0, 3) and t.cursor == [0, 0]


def test_bake_copies_glyph_cells():
    """text.rs:165-194."""
    a = _font()
    b = text.bake(b"\x01\x02\n\x05", a)
    assert b.shape == (48, 32, 3)
    assert (b[:24, :16] == a.get(1)).all() and (b[:24, 16:] == a.get(2)).all() and (b[24:, :16] == a.get(5)).all()
    assert (b[24:, 16:] == 0).all()


def test_hello_text_scene_renders_on_the_oracle(oracle):
    """hello.rs: 34 glyphs -> 68 triangles, none culled (face_cull = None), something is painted."""
    sc = scenes.hello_text()
    tgt = oracle.HostTarget(sc.w, sc.h, sc.fmt, True)
    st = oracle.render(sc.draws[0], tgt)
    assert st.prims.i == 68 and st.prims.o == 68 and st.frags.o > 10000
