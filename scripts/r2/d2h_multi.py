"""torchrun --nproc-per-node N scripts/r2/d2h_multi.py : what the platform gives when all ranks drain at once (VERDICT r1 item 4:
end-to-end efficiency 0.79 / 0.71 at 4 / 8 GPUs).  Every rank copies 1.2 GB device -> page-locked host memory (a) alone, one
rank after the other, (b) all ranks together; the aggregate of (b) over N x (a) is the ceiling of rn_sample's end-to-end
scaling -- that call is bound by this copy.  Also prints where each GPU hangs (NUMA node, PCIe bus)."""
import os, sys, time
import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 151552 * 100 * 10
d = torch.randn(N, dtype=torch.float64, device="cuda")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from rainier_b200 import api
pin = api.PinnedBuffer((N,), device=local)  # the library's own allocator (NUMA-placed next to the GPU)
h = torch.from_numpy(pin.array)
p = torch.cuda.get_device_properties(local)
pci = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
try:
    node = open("/sys/bus/pci/devices/%s/numa_node" % pci).read().strip()
except Exception:
    node = "?"


def copy_once():
    torch.cuda.synchronize()
    t = time.perf_counter()
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t


for _ in range(2):
    copy_once()
alone = None
for r in range(world):  # (a) one rank at a time
    if world > 1:
        dist.barrier()
    if r == rank:
        alone = min(copy_once() for _ in range(3))
if world > 1:
    dist.barrier()
together = []
for _ in range(5):  # (b) all ranks at once
    if world > 1:
        dist.barrier()
    together.append(copy_once())
tg = sorted(together)[len(together) // 2]
t = torch.tensor([alone, tg], dtype=torch.float64, device="cuda")
g = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(g, t)
else:
    g = [t]
print("rank %d gpu %s numa node %s: alone %.1f GB/s, all together %.1f GB/s" % (rank, pci, node, N * 8 / alone / 1e9, N * 8 / tg / 1e9), flush=True)
if world > 1:
    dist.barrier()
if rank == 0:
    a = sum(N * 8 / float(x[0]) / 1e9 for x in g)
    b = world * N * 8 / max(float(x[1]) for x in g) / 1e9
    print("aggregate: sum of alone %.1f GB/s, all together (max over ranks) %.1f GB/s -> ceiling of end-to-end scaling %.2f" % (a, b, b / a), flush=True)
if world > 1:
    dist.destroy_process_group()
