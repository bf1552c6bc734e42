"""The kernel SOURCE of retrofire_b200/csrc, executed on the CPU by the SIMT emulation in tests/emu and compared with
the oracle — the same test bodies as tests/test_gpu_{1_configs,2_api,3_adversarial}.py, so the warp-level logic (dependency rounds, warp-aggregated
allocation, bin sort, checkpoints, fast paths, error reporting, arena growth and replay) is checked in a container
without a GPU. This says nothing about the SASS or about speed: the parity tests proper are the `-m gpu` ones, which run
the nvcc-built library on a B200. The emulation library lives in tests/emu/_build, is loaded only here, and is unknown to
the retrofire_b200 package.
"""
import ctypes as C
import os
import sys

import pytest

import retrofire_b200 as rf
from retrofire_b200 import _ffi
from tests import test_gpu_1_configs as G1, test_gpu_2_api as G2, test_gpu_3_adversarial as G3

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402


@pytest.fixture(scope="module")
def device():
    """An rf_ctx of the emulation build. `_ffi._lib` is swapped only while the Device binds its library handle."""
    lib = C.CDLL(build_emu.build())
    for name, (res, args) in _ffi.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    saved = _ffi._lib
    _ffi._lib = lib
    try:
        dev = rf.Device(0)
    finally:
        _ffi._lib = saved
    assert dev.lib is lib
    yield dev
    dev.close()


# Everything in the three GPU parity files except the full-size scenes (minutes under emulation; they stay GPU-only) and, unless
# RF_EMU_FULL=1, the three cases that take more than 30 s each here (all 61 pass: 3.5 min).
SKIP = {
    "test_bunny_x16", "test_crates_1089_full_4k", "test_sprites_10k_full", "test_small_tris_1m_8k",
    "test_small_tris_8k_row_bands_cover_the_frame", "test_fuzz_random_frames_through_one_context",
    "test_closed_device_handles_are_not_reused",  # creates its own Device: needs the real library and a GPU
    "test_overflowing_pass_after_a_small_one_in_a_fresh_context",  # likewise (and a race between blocks is not reproducible here)
}
if os.environ.get("RF_EMU_FULL") != "1":
    SKIP |= {"test_arena_growth_replays_the_pass", "test_every_tile_heaviest_and_repeated_passes", "test_page_locked_geometry_is_dmad_directly"}
for _G in (G1, G2, G3):
    for _name in dir(_G):
        if _name.startswith("test_") and _name not in SKIP:
            globals()[_name] = getattr(_G, _name)


def test_fuzz_first_frames_under_emulation(device, oracle, monkeypatch):
    """The first 20 frames of the seeded fuzz (the GPU suite runs all 100); scratch/emu_fuzz.sh runs other seeds."""
    if "RF_FUZZ_FRAMES" not in os.environ:
        monkeypatch.setenv("RF_FUZZ_FRAMES", "20")
    G3.test_fuzz_random_frames_through_one_context(device, oracle)
