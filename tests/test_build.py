"""CPU-only checks of the native build: the library exists, exports the whole ABI, and the
coverage/depth kernels contain no fused multiply-adds (SURVEY §7 hard part 3)."""
import os
import re
import subprocess

import pytest

from retrofire_b200 import _ffi
from retrofire_b200 import build as rfbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    rfbuild.build()
    lib = _ffi.load()
    header = open(os.path.join(ROOT, "include", "retrofire_b200.h")).read()
    declared = set(re.findall(r"^(?:rf_status|void\*?|const char\*|uint32_t)\s+(rf_[a-z_0-9]+)\s*\(", header, re.M))
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rf_abi_version() == 2


def test_rust_sys_crate_and_ctypes_structs_follow_the_header():
    """The Rust -sys crate cannot be compiled here (no cargo): at least every header symbol must be declared in it, and the
    ctypes mirrors of rf_draw / rf_stats must have the header's field order."""
    header = open(os.path.join(ROOT, "include", "retrofire_b200.h")).read()
    rust = open(os.path.join(ROOT, "rust", "retrofire-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"^(?:rf_status|void\*?|const char\*|uint32_t)\s+(rf_[a-z_0-9]+)\s*\(", header, re.M))
    assert all(f"fn {name}(" in rust for name in declared), [n for n in declared if f"fn {n}(" not in rust]

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            for part in decl.split(","):
                m = re.search(r"(\w+)\s*(?:\[[^\]]*\])?\s*$", part.strip())
                if m and part.strip():
                    names.append(m.group(1))
        return names

    assert fields("rf_draw") == [f[0] for f in _ffi.RfDraw._fields_]
    assert fields("rf_stats") == [f[0] for f in _ffi.RfStats._fields_]
    for name in fields("rf_draw") + fields("rf_stats"):
        assert re.search(r"pub %s:" % name, rust), name


def test_rust_safe_shim_names_every_catalogue_pair_with_the_headers_ids():
    """rust/retrofire-b200 (not compilable here) must offer one shader value per catalogue fragment shader, with the numeric
    ids of the header's rf_vs_id / rf_fs_id enums, and set every rf_draw field."""
    header = open(os.path.join(ROOT, "include", "retrofire_b200.h")).read()
    shim = open(os.path.join(ROOT, "rust", "retrofire-b200", "src", "lib.rs")).read()
    ids = {name: int(val) for name, val in re.findall(r"\b(RF_(?:VS|FS)_[A-Z0-9_]+)\s*=\s*(\d+)", header)}
    fs_ids = {v for k, v in ids.items() if k.startswith("RF_FS_")}
    vs_ids = {v for k, v in ids.items() if k.startswith("RF_VS_")}
    used_fs = {int(v) for v in re.findall(r"const FS: u32 = (\d+);", shim)}
    used_vs = {int(v) for v in re.findall(r"const VS: u32 = (\d+);", shim)}
    for vs, fs in re.findall(r"mvp_shader!\(.*?(\d+),\s*(\d+)\);", shim, re.S):
        used_vs.add(int(vs)); used_fs.add(int(fs))
    # ids written symbolically (`sys::RF_FS_TEX_ONCE`) resolve through the -sys crate, whose constants must equal the header's
    rust_consts = {n: int(v) for n, v in re.findall(r"pub const (RF_(?:VS|FS)_[A-Z0-9_]+): u32 = (\d+);", open(os.path.join(ROOT, "rust", "retrofire-b200-sys", "src", "lib.rs")).read())}
    assert rust_consts == ids, set(rust_consts.items()) ^ set(ids.items())
    used_fs |= {rust_consts[n] for n in re.findall(r"const FS: u32 = sys::(RF_FS_\w+);", shim)}
    used_vs |= {rust_consts[n] for n in re.findall(r"const VS: u32 = sys::(RF_VS_\w+);", shim)}
    assert used_fs == fs_ids, used_fs ^ fs_ids
    assert used_vs == vs_ids, used_vs ^ vs_ids
    for name, val in re.findall(r"const (?:VS|FS): u32 = (\d+); // (RF_\w+)", shim):
        assert ids[val] == int(name), (val, name)   # the id written next to a symbolic comment is that symbol's value
    rust_sys = open(os.path.join(ROOT, "rust", "retrofire-b200-sys", "src", "lib.rs")).read()
    draw_fields = re.findall(r"pub (\w+):", re.search(r"pub struct rf_draw \{(.*?)\n\}", rust_sys, re.S).group(1))
    literal = re.search(r"let draw = sys::rf_draw \{(.*?)\n    \};", shim, re.S).group(1)
    for f in draw_fields:
        assert re.search(r"\b%s:" % f, literal), f


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import retrofire_b200 as rf
    with pytest.raises(rf.RetrofireError):
        rf.Device(0)


def test_ptx_has_no_fma_on_coverage_and_depth_chains(tmp_path):
    """User arithmetic must not be contracted: in PTX, fma.rn.f32 may appear only inside the powf
    expansion (19 per call) of the sRGB shaders; k_setup, k_edge_ckpt, k_walk and k_ckpt (setup, edge and span walks) have none."""
    ptx = open(rfbuild.ptx(str(tmp_path / "rf.ptx"))).read()
    counts, cur = {}, None
    for line in ptx.splitlines():
        m = re.search(r"\.entry\s+(\w+)", line)
        if m:
            cur = m.group(1)
        if "fma.rn.f32" in line and cur:
            counts[cur] = counts.get(cur, 0) + 1
    for name, n in counts.items():
        assert all(k not in name for k in ("k_assemble", "k_setup", "k_walk", "k_ckpt")), (name, n)
        if "k_raster" in name:
            assert n % (3 * 19) == 0, (name, n)   # powf(c, 1/2.2) x3 in FS_COLOR3F_SRGB, once per inlined copy
    assert "--use_fast_math" not in " ".join(rfbuild.NVCC_FLAGS)
    assert "-fmad=false" in rfbuild.NVCC_FLAGS


def test_sass_is_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def _compile_cpp_smoke(out):
    src = os.path.join(ROOT, "tests", "cpp_host_smoke.cpp")
    libdir = os.path.join(ROOT, "retrofire_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-o", out, src, f"-L{libdir}", "-lrf_b200", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.run(cmd, check=True)
    return out


def test_cpp_host_mirror_compiles_against_the_c_abi(tmp_path):
    """include/retrofire_b200.hpp (the compiled-language host mirror) builds and links with librf_b200.so."""
    rfbuild.build()
    assert os.path.exists(_compile_cpp_smoke(str(tmp_path / "cpp_host_smoke")))


@pytest.mark.gpu
def test_cpp_host_mirror_renders_hello_tri(tmp_path):
    """hello_tri through re::render(): centre pixel (114,102,128,255), 51,200 fragments (hello_tri.rs:47-53)."""
    exe = _compile_cpp_smoke(str(tmp_path / "cpp_host_smoke"))
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_header_is_plain_c(tmp_path):
    """include/retrofire_b200.h compiles as C99 with -pedantic -Werror (tests/c_header_check.c): what cgo, the Rust -sys crate or a
    C host bind is a C header, not a C++ one."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out = tmp_path / "c_header_check.o"
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(os.path.dirname(here), "include"), "-c",
                        os.path.join(here, "c_header_check.c"), "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_hot_kernels_do_not_spill():
    """Register budget of the hot kernels as built (cuobjdump --dump-resource-usage of the shipped library). k_raster<3> is capped at
    80 registers by its launch bounds (6 blocks per SM): one innocent-looking atomic in a fragment shader once pushed it from 48 to 378
    bytes of spills and the bunny step from 2.50 to 3.13 ms (profiles/r02_ab_tma_depth.txt) — this keeps such a change from going unseen."""
    rfbuild.build()
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    usage = {}
    for fn, res in re.findall(r"Function (\w+):\s*\n\s*(REG:[^\n]*)", out):
        usage[fn] = {k: int(v) for k, v in re.findall(r"(\w+):(\d+)", res)}
    ras = usage["_Z8k_rasterILi3ELb0EEv10PassParams"]
    assert ras["REG"] <= 80 and ras["STACK"] <= 64, ras
    asm_ = usage["_Z10k_assembleILi3ELb1EEv10PassParams"]
    assert asm_["REG"] <= 128, asm_
