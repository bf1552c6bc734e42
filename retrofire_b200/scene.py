"""Host-side mirror of core/src/render/scene.rs: `BBox` and `Obj` (geometry + bounding box + model-to-world transform).

The visibility test itself (`BBox::visibility`, scene.rs:81-87) runs on the device: pass `obj.bbox` as `DrawCall.bbox`
(rf_draw.bbox_cull) and hidden objects are skipped there. Pure host logic."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import mathx as mx

f32 = np.float32


@dataclass
class BBox:
    """scene.rs:22, 36-70. Default: the empty box (+inf, -inf) (scene.rs:112-117)."""
    low: np.ndarray = field(default_factory=lambda: np.full(3, np.inf, f32))
    upp: np.ndarray = field(default_factory=lambda: np.full(3, -np.inf, f32))

    def __post_init__(self):
        self.low, self.upp = np.asarray(self.low, f32).copy(), np.asarray(self.upp, f32).copy()

    @staticmethod
    def of(verts: np.ndarray) -> "BBox":
        """BBox::of (scene.rs:37-39): the box of the vertex positions (first three columns)."""
        b = BBox()
        for p in np.asarray(verts, f32)[:, :3]:
            b.extend(p)
        return b

    def extend(self, pt) -> None:
        """scene.rs:42-46: enlarge so that `pt` is just contained (f32::min / f32::max per component)."""
        pt = np.asarray(pt, f32)
        self.low, self.upp = np.fmin(self.low, pt), np.fmax(self.upp, pt)

    def is_empty(self) -> bool:
        """scene.rs:48-51: any low[i] >= upp[i]."""
        return bool((self.low >= self.upp).any())

    def contains(self, pt) -> bool:
        """scene.rs:54-57: low <= pt <= upp in every component."""
        pt = np.asarray(pt, f32)
        return bool(((self.low <= pt) & (pt <= self.upp)).all())

    def verts(self) -> np.ndarray:
        """scene.rs:60-69: the 8 corners, x-major."""
        (x0, y0, z0), (x1, y1, z1) = self.low, self.upp
        return np.array([[x, y, z] for x in (x0, x1) for y in (y0, y1) for z in (z0, z1)], f32)

    def as_array(self) -> np.ndarray:
        """(2, 3) float32 for `DrawCall.bbox`."""
        return np.stack([self.low, self.upp]).astype(f32)

    def __eq__(self, other) -> bool:
        return isinstance(other, BBox) and np.array_equal(self.low, other.low) and np.array_equal(self.upp, other.upp)


@dataclass
class Obj:
    """scene.rs:13-34: `Obj::new` / `with_transform` compute the box from the mesh."""
    faces: np.ndarray
    verts: np.ndarray
    tf: np.ndarray = field(default_factory=mx.identity)
    bbox: BBox = None

    def __post_init__(self):
        if self.bbox is None:
            self.bbox = BBox.of(self.verts)
