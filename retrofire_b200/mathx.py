"""Host-side f32 matrix helpers mirroring retrofire-core's `math::mat` constructors.

These run once per draw on the host (cam.rs / mat.rs are callers of the hot
path, not part of it) but the golden images of the reference pin the *exact* f32 values of
`perspective`, `viewport` and `then`, so the operation order below follows
core/src/math/mat.rs literally:

* `dot`  — left fold from 0.0                           (math/vec.rs:231-238)
* `then`/`compose` — dot(lhs row j, rhs column i)       (math/mat.rs:268-299)
* `perspective`, `orthographic`, `viewport`             (math/mat.rs:1255-1315)
* `translate3`, `scale3`, `rotate_x/y/z`                (math/mat.rs:1035-1193)

All matrices are row-major 4x4 `numpy.float32` arrays.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def _dot(a, b) -> np.float32:
    acc = f32(0.0)
    for x, y in zip(a, b):
        acc = f32(acc + f32(f32(x) * f32(y)))
    return acc


def mat(rows) -> np.ndarray:
    return np.asarray(rows, dtype=np.float32).reshape(4, 4)


def identity() -> np.ndarray:
    return np.eye(4, dtype=np.float32)


def compose(outer: np.ndarray, inner: np.ndarray) -> np.ndarray:
    """`outer.compose(inner)`: apply `inner` first, then `outer` (mat.rs:268-286)."""
    out = np.empty((4, 4), dtype=np.float32)
    it = inner.T
    for j in range(4):
        for i in range(4):
            out[j, i] = _dot(outer[j], it[i])
    return out


def then(first: np.ndarray, second: np.ndarray) -> np.ndarray:
    """`first.then(&second)` == `second.compose(&first)` (mat.rs:287-299)."""
    return compose(second, first)


def translate3(x, y, z) -> np.ndarray:
    return mat([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z], [0, 0, 0, 1]])


def scale3(x, y, z) -> np.ndarray:
    return mat([[x, 0, 0, 0], [0, y, 0, 0], [0, 0, z, 0], [0, 0, 0, 1]])


def _sincos(a):
    a = f32(a)
    return f32(math.sin(float(a))), f32(math.cos(float(a)))


def rotate_x(a) -> np.ndarray:
    s, c = _sincos(a)
    return mat([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]])


def rotate_y(a) -> np.ndarray:
    s, c = _sincos(a)
    return mat([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]])


def rotate_z(a) -> np.ndarray:
    s, c = _sincos(a)
    return mat([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])


def perspective(focal_ratio, aspect_ratio, near, far) -> np.ndarray:
    """mat.rs:1255-1281."""
    focal_ratio, aspect_ratio, near, far = f32(focal_ratio), f32(aspect_ratio), f32(near), f32(far)
    assert focal_ratio > 0 and aspect_ratio > 0 and near > 0 and far > near
    e00 = focal_ratio
    e11 = f32(e00 * aspect_ratio)
    e22 = f32(f32(far + near) / f32(far - near))
    e23 = f32(f32(f32(f32(2.0) * far) * near) / f32(near - far))
    return mat([[e00, 0, 0, 0], [0, e11, 0, 0], [0, 0, e22, e23], [0, 0, 1, 0]])


def orthographic(lbn, rtf) -> np.ndarray:
    """mat.rs:1283-1299."""
    x0, y0, z0 = (f32(v) for v in lbn)
    x1, y1, z1 = (f32(v) for v in rtf)
    dx, dy, dz = f32(f32(x1 - x0) / f32(2)), f32(f32(y1 - y0) / f32(2)), f32(f32(z1 - z0) / f32(2))
    cx, cy, cz = f32(x0 + dx), f32(y0 + dy), f32(z0 + dz)
    idx, idy, idz = f32(f32(1) / dx), f32(f32(1) / dy), f32(f32(1) / dz)
    return mat([[idx, 0, 0, f32(-cx * idx)], [0, idy, 0, f32(-cy * idy)], [0, 0, idz, f32(-cz * idz)], [0, 0, 0, 1]])


def viewport(start, end) -> np.ndarray:
    """`viewport(pt2(x0,y0)..pt2(x1,y1))`, mat.rs:1304-1315."""
    x0, y0 = f32(start[0]), f32(start[1])
    x1, y1 = f32(end[0]), f32(end[1])
    dx, dy = f32(f32(x1 - x0) / f32(2)), f32(f32(y1 - y0) / f32(2))
    return mat([[dx, 0, 0, f32(x0 + dx)], [0, dy, 0, f32(y0 + dy)], [0, 0, 1, 0], [0, 0, 0, 1]])


def fov_equiv35mm(mm) -> np.float32:
    """cam.rs:128: mm / (36/2)."""
    return f32(f32(mm) / f32(18.0))


def fov_diagonal(angle_rad, aspect) -> np.float32:
    """cam.rs:134-141."""
    aspect = f32(aspect)
    ratio = f32(f32(1.0) / f32(math.tan(float(f32(f32(angle_rad) / f32(2.0))))))
    diag = f32(math.sqrt(float(f32(f32(1.0) + f32(f32(f32(1.0) / aspect) / aspect)))))
    return f32(ratio * diag)


def normalize(v) -> np.ndarray:
    v = np.asarray(v, dtype=np.float32)
    return (v * f32(1.0 / math.sqrt(float(_dot(v, v))))).astype(np.float32)
