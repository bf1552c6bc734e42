"""Binary PPM (P6) read/write — the reference's only persistence format (core/src/util/pnm.rs:281-296)."""
from __future__ import annotations

import gzip

import numpy as np


def read_ppm(path: str) -> np.ndarray:
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as fh:
        data = fh.read()
    return parse_ppm(data)


def parse_ppm(data: bytes) -> np.ndarray:
    """Returns (h, w, 3) uint8. Header: P6 <ws> w <ws> h <ws> maxval <single ws> raster; '#' comments allowed."""
    pos, toks = 0, []
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] not in (b"\n", b""):
                pos += 1
            continue
        start = pos
        while not data[pos:pos + 1].isspace():
            pos += 1
        toks.append(data[start:pos])
    pos += 1
    assert toks[0] == b"P6" and int(toks[3]) == 255, toks
    w, h = int(toks[1]), int(toks[2])
    return np.frombuffer(data, dtype=np.uint8, count=w * h * 3, offset=pos).reshape(h, w, 3).copy()


def save_ppm(path: str, rgb: np.ndarray) -> None:
    rgb = np.ascontiguousarray(rgb[..., :3], dtype=np.uint8)
    h, w = rgb.shape[:2]
    with open(path, "wb") as fh:
        fh.write(b"P6 %d %d 255\n" % (w, h))
        fh.write(rgb.tobytes())
