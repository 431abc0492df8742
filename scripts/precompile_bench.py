"""Emit + NVRTC-compile (no device needed) the kernels bench.py's cfg3 / cfg5 side measurements will ask for, into
$RN_KERNEL_CACHE, so that the GPU box loads cubins instead of compiling for a minute per rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from rainier_b200 import api

N_STEPS = 5
for name, cfg in (("cfg3_primal", api.make_config(iterations=2, warmupIterations=0, sampler=api.HMCSampler(N_STEPS), stepSizeTuner=api.StaticStepSize(0.01),
                                                  massMatrixTuner=api.IdentityMassMatrixTuner(), launchIterations=2)),
                  ("cfg5_primal", api.make_config(iterations=2, warmupIterations=30, sampler=api.HMCSampler(N_STEPS), stepSizeTuner=api.DualAvgTuner(0.8),
                                                  massMatrixTuner=api.IdentityMassMatrixTuner(), launchIterations=2))):
    f = os.path.join(ROOT, "build", "models", name + ".npz")
    if not os.path.exists(f):
        continue
    z = np.load(f)
    m = api.CudaModel(z["rir"].tobytes(), [z["c%d" % i] for i in range(int(z["ncols"]))], device=-1)
    m.emit_cubin(cfg)
    m.close()
    print("precompiled", name)
