"""TEST INFRASTRUCTURE ONLY: compiles retrofire_b200/csrc for the HOST against tests/emu/include (a SIMT emulation of
the CUDA constructs the kernels use), into tests/emu/_build/librf_b200_emu.so.

The kernel sources are used as they are, apart from four mechanical rewrites that g++ needs (each asserted to apply):
  * `kernel<<<grid, block[, smem[, stream]]>>>(args);`  ->  `emu::launch("kernel", grid, block, ..., [&] { kernel(args); });`
  * `extern __shared__ T name[];`                        ->  `T* name = (T*)emu::dyn_smem();`
  * `__shared__ T name...;`                              ->  `static T name...;`   (blocks run one after the other)
  * the three inline-PTX sites (laneid, st.release / ld.acquire of the peer barrier) -> their C++ meaning;
    WarpSmem's ld/st.shared PTX is switched off with the source's own -DRF_SMEM_ASM=0.

Nothing in retrofire_b200/ knows about this library; only tests/test_emu_kernels.py loads it.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "retrofire_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
OUT = os.path.join(BUILD, "librf_b200_emu.so")
FILES = ["rf_api.cu", "rf_device.cuh", "rf_geometry.cuh", "rf_raster.cuh", "rf_order.cuh", "rf_peer.cuh"]

LAUNCH = re.compile(r"(?P<k>\b\w+(?:<[^<>;()]*>)?)\s*<<<(?P<cfg>[^;]*?)>>>\s*\((?P<args>[^;]*)\);")


def transform(name: str, text: str) -> str:
    n_launch = len(re.findall(r"<<<", text))
    text, n = LAUNCH.subn(lambda m: f"emu::launch(\"{m.group('k')}\", {m.group('cfg')}, [&] {{ {m.group('k')}({m.group('args')}); }});", text)
    assert n == n_launch, f"{name}: {n_launch} kernel launches, {n} rewritten"
    text, n_dyn = re.subn(r"extern\s+__shared__\s+(?P<t>[\w ]+?)\s+(?P<n>\w+)\[\];", lambda m: f"{m.group('t')}* {m.group('n')} = ({m.group('t')}*)emu::dyn_smem();", text)
    text, n_sta = re.subn(r"(?m)^(\s*)__shared__\s+", r"\1static ", text)
    assert "__shared__" not in re.sub(r"//.*", "", text), f"{name}: a __shared__ declaration was not rewritten"
    asm_sites = {
        'asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));': "l = threadIdx.x & 31u;",
        'asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");': "*(volatile uint32_t*)theirs = epoch;",
        'asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");': "v = *(const volatile uint32_t*)mine;",
    }
    for pat, rep in asm_sites.items():
        text = text.replace(pat, rep)
    text = text.replace("__noinline__", "__attribute__((noinline))")
    return text


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in FILES] + [os.path.join(ROOT, "include", "retrofire_b200.h"), os.path.abspath(__file__),
                                                     os.path.join(HERE, "include", "cuda_runtime.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, defines: dict | None = None, out: str | None = None) -> str:
    out = out or OUT
    if os.environ.get("RF_EMU_CXXFLAGS") and out == OUT:  # debugging builds (sanitizers, -O0 ...) never replace the default library
        out = OUT.replace(".so", "_custom.so")
    if not (force or defines or out != OUT or stale() or os.environ.get("RF_EMU_CXXFLAGS")):
        return out
    src = os.path.join(BUILD, "src", "retrofire_b200", "csrc")  # same depth as the original: "../../include/retrofire_b200.h" resolves
    os.makedirs(src, exist_ok=True)
    os.makedirs(os.path.join(BUILD, "src", "include"), exist_ok=True)
    with open(os.path.join(ROOT, "include", "retrofire_b200.h")) as f, open(os.path.join(BUILD, "src", "include", "retrofire_b200.h"), "w") as g:
        g.write(f.read())
    for name in FILES:
        with open(os.path.join(CSRC, name)) as f:
            text = transform(name, f.read())
        with open(os.path.join(src, name + (".cpp" if name.endswith(".cu") else "")), "w") as g:
            g.write(text)
    # the remaining inline PTX (WarpSmem) must be compiled out by RF_SMEM_ASM=0; any other asm would be a new site to handle
    left = [n for n in FILES if "asm volatile" in re.sub(r"(?s)#if RF_SMEM_ASM.*?#else", "", open(os.path.join(src, n + (".cpp" if n.endswith(".cu") else ""))).read())]
    assert not left, f"unhandled inline PTX in {left}"
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing", "-fPIC", "-shared", "-w", "-x", "c++",
           "-DRF_SMEM_ASM=0", f"-I{os.path.join(HERE, 'include')}"] + [f"-D{k}={v}" for k, v in (defines or {}).items()] + \
          os.environ.get("RF_EMU_CXXFLAGS", "").split() + ["-o", out, os.path.join(src, "rf_api.cu.cpp")]
    subprocess.run(cmd, check=True, cwd=src)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
