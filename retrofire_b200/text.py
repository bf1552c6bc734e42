"""Text as texture-mapped geometry: host-side mirror of core/src/render/text.rs and tex.rs `Atlas`.

One quad (two triangles) per glyph, texture coordinates into a grid atlas; the geometry is then an ordinary
render() call with the FS_TEX_CLAMP fragment shader (`Text::sample` = SamplerClamp, text.rs:59-62). Pure host logic:
nothing here touches the GPU or the oracle."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .api import Texture

f32 = np.float32


class Atlas:
    """tex.rs:39-50, 141-186 — `Layout::Grid { sub_dims }` over one texture."""

    def __init__(self, sub_dims: Tuple[int, int], texture: Texture):
        self.sub_dims = (int(sub_dims[0]), int(sub_dims[1]))
        self.texture = texture

    def rect(self, i: int):
        """Top-left and bottom-right pixel coordinates of sub-texture `i` (tex.rs:149-158; u32 arithmetic)."""
        sw, sh = self.sub_dims
        per_row = self.texture.w // sw
        x0, y0 = i % per_row * sw, i // per_row * sh
        return (x0, y0), (x0 + sw, y0 + sh)

    def get(self, i: int) -> np.ndarray:
        """The texels of sub-texture `i` (tex.rs:165-168)."""
        (x0, y0), (x1, y1) = self.rect(i)
        assert x1 <= self.texture.w and y1 <= self.texture.h, "glyph index out of bounds"
        return self.texture.data[y0:y1, x0:x1]

    def coords(self, i: int) -> np.ndarray:
        """uv of the top-left, top-right, bottom-left, bottom-right corners (tex.rs:178-185): pixel / f32 dims in f32."""
        (px0, py0), (px1, py1) = self.rect(i)
        tw, th = f32(self.texture.w), f32(self.texture.h)
        x0, y0, x1, y1 = f32(px0) / tw, f32(py0) / th, f32(px1) / tw, f32(py1) / th
        return np.array([[x0, y0], [x1, y0], [x0, y1], [x1, y1]], dtype=f32)


class Text:
    """text.rs:12-95: glyph quads appended at a cursor; '\\n' moves the cursor to the start of the next row."""

    def __init__(self, font: Atlas):
        self.font = font
        self.verts: List[List[float]] = []   # [x, y, z, u, v]
        self.faces: List[List[int]] = []
        self.cursor = [f32(0), f32(0)]

    def clear(self) -> None:
        self.cursor = [f32(0), f32(0)]
        self.verts.clear()
        self.faces.clear()

    def _write_char(self, idx: int) -> None:     # text.rs:64-94
        gw, gh = f32(self.font.sub_dims[0]), f32(self.font.sub_dims[1])
        tl, tr, bl, br = self.font.coords(idx)
        x, y = self.cursor
        n = len(self.verts)
        self.verts += [[x, y, f32(0), *tl], [x + gw, y, f32(0), *tr], [x, y + gh, f32(0), *bl], [x + gw, y + gh, f32(0), *br]]
        self.faces += [[n, n + 1, n + 3], [n, n + 3, n + 2]]
        self.cursor[0] = x + gw

    def write(self, s) -> "Text":
        """io::Write (bytes: one glyph per byte, text.rs:103-136) or fmt::Write (str: one glyph per char, :138-162)."""
        gh = f32(self.font.sub_dims[1])
        for ch in (s if isinstance(s, (bytes, bytearray)) else map(ord, s)):
            if ch == 10:
                self.cursor = [f32(0), self.cursor[1] + gh]
            else:
                self._write_char(int(ch))
        return self

    @property
    def geom(self):
        """(faces (n,3) uint32, verts (m,5) float32) — `Mesh<TexCoord>`."""
        return np.array(self.faces, dtype=np.uint32).reshape(-1, 3), np.array(self.verts, dtype=f32).reshape(-1, 5)


def bake(s: bytes, font: Atlas) -> np.ndarray:
    """text.rs:165-194: copies glyph cells into a (rows*gh, cols*gw, C) buffer; default (zero) texels elsewhere."""
    rows = s.split(b"\n")
    n_rows, n_cols = len(rows), max(len(r) for r in rows)
    gw, gh = font.sub_dims
    if n_rows == 0 or n_cols == 0:
        return np.zeros((0, 0, font.texture.data.shape[2]), np.uint8)
    buf = np.zeros((n_rows * gh, n_cols * gw, font.texture.data.shape[2]), np.uint8)
    for r, row in enumerate(rows):
        for c, ch in enumerate(row):
            buf[r * gh:(r + 1) * gh, c * gw:(c + 1) * gw] = font.get(ch)
    return buf
