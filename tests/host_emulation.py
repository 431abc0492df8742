"""
Debug aid for the CPU box (no GPU here): compiles the CUDA source the emitter produced with g++ under
-DRN_HOST_EMULATION (rn_prelude.cuh maps the few CUDA constructs used to plain C++) and runs the emitted
rn_density() on the host, one "thread" at a time.  This checks the *emitter* (lowering, reverse-mode adjoints,
accumulation order) against the oracle without a device.  It is test infrastructure only -- the product never
compiles or runs this; the product path fails loudly without CUDA.
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_SHIM = r"""
extern "C" void emu_density(const double* q, int chains, double* out, const double* data, int* err) {
  blockDim.x = 1; gridDim.x = (unsigned)chains; threadIdx.x = 0;
  for (int c = 0; c < chains; c++) { blockIdx.x = (unsigned)c; rn_k_density(q, out, data, err, chains); }
}
"""


def compile_source(src, fast=False):
    d = os.path.join(tempfile.gettempdir(), "rn_emul")
    os.makedirs(d, exist_ok=True)
    key = hashlib.sha1((src + str(fast)).encode()).hexdigest()[:16]
    so = os.path.join(d, key + ".so")
    if not os.path.exists(so):
        cpp = os.path.join(d, key + ".cpp")
        with open(cpp, "w") as f:
            f.write(src + _SHIM)
        flags = ["-O1", "-std=c++17", "-fPIC", "-shared", "-DRN_HOST_EMULATION", "-w"]
        flags.append("-ffp-contract=fast" if fast else "-ffp-contract=off")
        subprocess.run(["g++"] + flags + [cpp, "-o", so], check=True)
    return C.CDLL(so)


def density(src, q, cols, model, fast=False):
    """q: [chains][n] -> [chains][n+1] using the emitted code.  `model`: the CudaModel the source came from (its
    rn_model_pack_columns lays the columns out exactly as rn_model_create uploads them: tile-major per target)."""
    L = compile_source(src, fast)
    q = np.ascontiguousarray(q, dtype=np.float64)
    chains, n = q.shape
    qt = np.ascontiguousarray(q.T)
    out = np.zeros((n + 1, chains))
    data = model.pack_columns()
    err = C.c_int(0)
    L.emu_density(C.c_void_p(qt.ctypes.data), chains, C.c_void_p(out.ctypes.data), C.c_void_p(data.ctypes.data), C.byref(err))
    return np.ascontiguousarray(out.T), err.value
