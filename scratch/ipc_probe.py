"""Feasibility: CUDA IPC between torchrun ranks + P2P stores over NVLink from a trivial kernel (torch ops on an imported pointer)."""
import ctypes as C, os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = C.CDLL("libcudart.so.12")
n = 64 << 20
p = C.c_void_p()
assert rt.cudaMalloc(C.byref(p), C.c_size_t(n)) == 0
class H(C.Structure):
    _fields_ = [("r", C.c_ubyte * 64)]
h = H()
assert rt.cudaIpcGetMemHandle(C.byref(h), p) == 0
handles = [None] * world
dist.all_gather_object(handles, bytes(h))
peers = []
for r in range(world):
    if r == rank:
        peers.append(p.value); continue
    q = C.c_void_p()
    hh = H.from_buffer_copy(handles[r])
    rt.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), H, C.c_uint]
    rc = rt.cudaIpcOpenMemHandle(C.byref(q), hh, 1)
    print(rank, "open handle of", r, "rc", rc, hex(q.value or 0), flush=True)
    assert rc == 0
    peers.append(q.value)
# wrap the peer pointer as a torch tensor through __cuda_array_interface__
class Wrap:
    def __init__(self, ptr, n): self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}
mine = torch.as_tensor(Wrap(peers[rank], n), device="cuda")
mine.fill_(rank + 1)
torch.cuda.synchronize(); dist.barrier()
other = torch.as_tensor(Wrap(peers[(rank + 1) % world], n), device="cuda")
src = torch.full((n,), 100 + rank, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(10): other[: n // 2].copy_(src[: n // 2])
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
dist.barrier()
print(rank, "peer store %.1f GB/s; my buffer now holds" % (n / 2 / dt / 1e9), int(mine[0]), int(mine[n - 1]), flush=True)
dist.barrier()
dist.destroy_process_group()
