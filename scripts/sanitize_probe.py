"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box: the TMA tile pipeline, the
K-warps-per-chain groups (named barriers, cross-warp scratch), the scatter path, the thread-per-chain kernels, the
pipelined rn_sample drain and the diagnostics reductions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle.rainier_py import configs
from rainier_b200 import abi, api


def static(it, nsteps, eps, **kw):
    return api.make_config(iterations=it, warmupIterations=0, sampler=api.HMCSampler(nsteps), stepSizeTuner=api.StaticStepSize(eps),
                           massMatrixTuner=api.IdentityMassMatrixTuner(), **kw)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "wpc"):
    rir, cols = configs.logreg(700, 4).compile(False)
    for k in ("1", "2"):
        os.environ["RN_WPC_K"] = k
        m = api.CudaModel(rir, cols)
        tr = m.sample(static(3, 2, 0.02, backend=abi.RN_BACKEND_WARP), seeds=np.arange(11) + 1)  # HMC: TMA tiles, ragged CTA
        tr2 = m.sample(api.SamplerConfig(iterations=3, warmupIterations=6, backend=abi.RN_BACKEND_WARP), seeds=np.arange(5) + 1)  # EHMC: ldg path
        print("wpc K=%s ok" % k, float(tr.chains.mean()), float(tr2.chains.mean()), flush=True)
        m.close()
    del os.environ["RN_WPC_K"]
    prir, pcols = configs.poisson_glm(40, 1300).compile(False)
    os.environ["RN_WPC_K"] = "2"
    m = api.CudaModel(prir, pcols)
    tr = m.sample(static(2, 2, 0.01, backend=abi.RN_BACKEND_WARP), seeds=np.arange(6) + 1)
    print("poisson scatter K=2 ok", float(tr.chains.mean()), flush=True)
    del os.environ["RN_WPC_K"]
if which in ("all", "tpc"):
    rir, cols = configs.eight_schools().compile(True)
    m = api.CudaModel(rir, cols)
    os.environ["RN_SAMPLE_BLOCKS"] = "3"
    tr = m.sample(api.SamplerConfig(iterations=20, warmupIterations=60), seeds=np.arange(300) + 1)
    del os.environ["RN_SAMPLE_BLOCKS"]
    print("tpc ok", float(tr.chains.mean()), flush=True)
    import torch
    s = api.CudaSampler(m, api.SamplerConfig(iterations=30, warmupIterations=40), seeds=np.arange(64) + 1)
    d = torch.empty((30, m.nVars, 64), dtype=torch.float64, device="cuda")
    s.warmup(-1)
    s.run(30, d.data_ptr())
    print("diag ok", s.diagnostics(d.data_ptr(), 30)[:2], flush=True)
    s.close()
