#!/bin/bash
# The seeded fuzz of tests/test_gpu_3_adversarial.py through the SIMT emulation of the kernel source (no GPU), with other seeds than the
# committed one: RF_FUZZ_SEED / RF_FUZZ_FRAMES select the stream. Usage: scratch/emu_fuzz.sh "<seeds>" <frames>
cd "$(dirname "$0")/.."
for seed in ${1:-7 8 9}; do
  RF_FUZZ_SEED=$seed RF_FUZZ_FRAMES=${2:-12} timeout ${3:-1500} python - <<'PY'
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests/emu")
import retrofire_b200 as rf
from retrofire_b200 import _ffi
from oracle import rfo
from tests import test_gpu_3_adversarial as G
import build_emu
t = time.time()
rfo.build(); rfo.load()
lib = C.CDLL(build_emu.build())
for name, (res, args) in _ffi.SYMBOLS.items():
    fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
saved, _ffi._lib = _ffi._lib, lib
try:
    dev = rf.Device(0)
finally:
    _ffi._lib = saved
G.test_fuzz_random_frames_through_one_context(dev, rfo)
print("seed", os.environ["RF_FUZZ_SEED"], "frames", os.environ["RF_FUZZ_FRAMES"], "bit-exact,", round(time.time() - t), "s", flush=True)
PY
  echo "seed $seed rc $?"
done
