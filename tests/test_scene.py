"""render/scene.rs unit tests (scene.rs:131-190), on the host-side mirror, with the reference's own vectors."""
import numpy as np

from retrofire_b200.scene import BBox, Obj


def test_bbox_default():
    assert BBox().is_empty()
    assert not BBox().contains([0, 0, 0])


def test_bbox_extend():
    b = BBox([-1.0, -2.0, -3.0], [5.0, 3.0, 2.0])
    b.extend([1.0, 1.0, 1.0])
    assert b == BBox([-1.0, -2.0, -3.0], [5.0, 3.0, 2.0])
    b.extend([-2.0, 3.0, 3.0])
    assert b == BBox([-2.0, -2.0, -3.0], [5.0, 3.0, 3.0])


def test_bbox_is_empty():
    assert BBox([-1.0, 0.0, -1.0], [1.0, 0.0, 1.0]).is_empty()
    assert BBox([-1.0, -1.0, 1.0], [1.0, 1.0, -1.0]).is_empty()
    assert not BBox([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]).is_empty()
    assert not BBox([-1.0, 10.0, -1.0], [1.0, np.inf, 1.0]).is_empty()


def test_bbox_contains():
    assert not BBox([-1.0, 0.0, -1.0], [1.0, 0.0, 1.0]).contains([0.0, 1.0, 0.0])
    assert BBox([-1.0, 0.0, -1.0], [1.0, 0.0, 1.0]).contains([0.0, 0.0, 0.0])
    assert BBox([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]).contains([-1.0, 0.0, 0.0])
    assert BBox([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]).contains([0.0, 0.0, 0.0])


def test_obj_computes_its_box_and_corner_order():
    """Obj::with_transform (scene.rs:29-33) and BBox::verts (scene.rs:60-69)."""
    verts = np.array([[0, 1, 2, 9], [-1, 4, 0, 9], [3, -2, 1, 9]], np.float32)      # extra columns are attributes
    o = Obj(np.array([[0, 1, 2]], np.uint32), verts)
    assert o.bbox == BBox([-1, -2, 0], [3, 4, 2])
    c = o.bbox.verts()
    assert c[0].tolist() == [-1, -2, 0] and c[1].tolist() == [-1, -2, 2] and c[4].tolist() == [3, -2, 0] and c[7].tolist() == [3, 4, 2]
    assert o.bbox.as_array().shape == (2, 3)
