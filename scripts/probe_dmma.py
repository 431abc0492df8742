#!/usr/bin/env python
"""
scripts/probe_dmma.py -- runs scripts/probes/dmma_probe.cu on a GPU box (measurement probe, not product):
checks the assumed operand mapping of mma.sync.m8n8k4.f64 on the sampler's own data layouts against a torch fp64 matmul
and reports the sustained rate of the tensor-core form vs the rows-across-lanes DFMA form of Z = X * B at cfg 3's shape.
Prints one JSON line.
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch
    so = os.path.join(ROOT, "build", "libdmma_probe.so")
    if not os.path.exists(so):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-shared",
                        "-Xcompiler", "-fPIC", os.path.join(ROOT, "scripts", "probes", "dmma_probe.cu"), "-o", so], check=True)
    L = C.CDLL(so)
    L.dmma_probe_run.restype = C.c_float
    L.dmma_probe_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    rows, d, chains = 100000, 50, 2048
    n_tiles, dpad = (rows + 31) // 32, 52
    g = torch.Generator(device="cuda").manual_seed(1)
    Xr = torch.zeros((n_tiles * 32, dpad), dtype=torch.float64, device="cuda")
    Xr[:rows, :d] = torch.randn((rows, d), dtype=torch.float64, device="cuda", generator=g)
    B = torch.zeros((dpad, chains), dtype=torch.float64, device="cuda")
    B[:d] = torch.randn((d, chains), dtype=torch.float64, device="cuda", generator=g)
    X = Xr.reshape(n_tiles, 32, dpad).permute(0, 2, 1).contiguous()  # tile-major [tile][column][32 rows]
    ref = Xr @ B
    out = {"shape": {"rows": rows, "d": d, "d_padded": dpad, "chains": chains}}
    flops = 2.0 * n_tiles * 32 * dpad * chains
    for which, name in ((0, "dmma_m8n8k4"), (1, "dfma_rows_across_lanes")):
        Z = torch.full((n_tiles * 32, chains), float("nan"), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ms = L.dmma_probe_run(which, X.data_ptr(), B.data_ptr(), Z.data_ptr(), n_tiles, dpad, chains, 10)
        torch.cuda.synchronize()
        err = float(((Z - ref).abs() / ref.abs().clamp_min(1e-9)).max())
        out[name] = {"ms": ms, "tflops": flops / (ms * 1e-3) / 1e12 if ms > 0 else None, "max_rel_err_vs_torch_matmul": err}
    out["mapping_ok"] = out["dmma_m8n8k4"]["max_rel_err_vs_torch_matmul"] < 1e-12
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
