"""Where does page-locked memory land relative to the GPU, and can this process steer it?  (GPU box only.)"""
import ctypes, glob, os, subprocess, sys, time
import torch
print(subprocess.run("lscpu | grep -i 'numa\\|socket\\|model name' ; nvidia-smi topo -m | head -6", shell=True, capture_output=True, text=True).stdout)
bus = torch.cuda.get_device_properties(0)
pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id) if hasattr(bus, "pci_bus_id") else None
print("gpu pci:", pci)
for f in ("numa_node", "local_cpulist"):
    try:
        print(f, open("/sys/bus/pci/devices/%s/%s" % (pci, f)).read().strip())
    except Exception as e:
        print(f, "unreadable:", e)
libc = ctypes.CDLL("libc.so.6", use_errno=True)
mask = (ctypes.c_ulong * 16)()
mask[0] = 1
rc = libc.syscall(238, 1, mask, 1025)  # set_mempolicy(MPOL_PREFERRED, node 0)
print("set_mempolicy rc", rc, "errno", ctypes.get_errno())
libc.syscall(238, 0, None, 0)
N = 1 << 27  # 1 GiB of doubles
d = torch.randn(N, dtype=torch.float64, device="cuda")
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes:", [os.path.basename(n) for n in nodes])
def cpus_of(node):
    s = open(node + "/cpulist").read().strip()
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out
full = os.sched_getaffinity(0)
for node in nodes:
    try:
        os.sched_setaffinity(0, set(cpus_of(node)) & full or full)
        h = torch.empty(N, dtype=torch.float64).pin_memory()
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        print("pinned alloc while on %s: D2H %.1f GB/s" % (os.path.basename(node), N * 8 / best / 1e9), flush=True)
        del h
    except Exception as e:
        print(node, "failed", e)
os.sched_setaffinity(0, full)
