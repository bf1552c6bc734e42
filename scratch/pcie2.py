import torch, time
n = 64 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(do_h2d, do_d2h, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if do_h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if do_d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for _ in range(2): run(True, True, 3)
a, b, c = run(True, False), run(False, True), run(True, True)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both: %.2f ms for 64MB each way -> H2D %.1f + D2H %.1f GB/s (serial would be %.2f ms)" % (n/a/1e9, n/b/1e9, c*1e3, n/c/1e9, n/c/1e9, (a+b)*1e3))
# small chunks like the bench: 8 x 1.9 MB H2D with a sync each, while a 66 MB D2H runs
import numpy as np
chunk = 1_900_000
def run2(reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s2): h2[:66_000_000].copy_(d2[:66_000_000], non_blocking=True)
        for k in range(8):
            with torch.cuda.stream(s1): d1[k*chunk:(k+1)*chunk].copy_(h1[k*chunk:(k+1)*chunk], non_blocking=True)
            s1.synchronize()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
print("bench-like step (66 MB D2H + 8 x 1.9 MB H2D with sync): %.2f ms" % (run2()*1e3))
import subprocess
print(subprocess.run(["nvidia-smi","--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max","--format=csv"],capture_output=True,text=True).stdout)
print(subprocess.run(["nvidia-smi","topo","-m"],capture_output=True,text=True).stdout[:1500])
