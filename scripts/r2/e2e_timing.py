"""RN_TIMING breakdown of rn_sample at the BASELINE chain counts: funnel HMC(5) x 100 iterations, page-locked
caller buffer, for 4096 / 8192 / 151552 chains; plus the summaries-only call (diagnostics, samples == NULL)."""
import ctypes as CT
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
os.environ["RN_TIMING"] = "1"
import numpy as np

from rainier_b200 import api

rir = open(os.path.join(os.path.dirname(__file__), "..", "..", "rainier_b200", "models", "funnel10.rir"), "rb").read()
model = api.CudaModel(rir, [], device=0)
I_ = 100
for C_ in (4096, 8192, 151552):
    cfg = api.make_config(iterations=I_, warmupIterations=0, sampler=api.HMCSampler(5), stepSizeTuner=api.StaticStepSize(0.1),
                          massMatrixTuner=api.IdentityMassMatrixTuner(), launchIterations=I_)
    c, keep = api.lower_config(cfg)
    seeds = np.arange(C_, dtype=np.int64)
    pin = api.PinnedBuffer((C_, I_, 10), device=0)
    for mode in ("samples", "diagnostics"):
        diag = np.empty((10, 2))
        c.diagnostics = diag.ctypes.data_as(CT.POINTER(CT.c_double)) if mode == "diagnostics" else None
        for k in range(5):
            sys.stderr.flush()
            t = time.perf_counter()
            rc = api.lib().rn_sample(model.h, CT.byref(c), seeds.ctypes.data, C_, pin.array.ctypes.data if mode == "samples" else None, None, None)
            dt = (time.perf_counter() - t) * 1e3
            print("chains %6d %-11s call %d rc %d %.2f ms" % (C_, mode, k, rc, dt), flush=True)
    pin.close()
