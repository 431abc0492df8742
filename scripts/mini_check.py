import sys, numpy as np
sys.path.insert(0, '.')
from rainier_b200 import api
from oracle.rainier_py.binding import OracleModel
rir = open('rainier_b200/models/funnel10.rir','rb').read()
cfg = api.make_config(iterations=20, warmupIterations=30, sampler=api.HMCSampler(5), stepSizeTuner=api.DualAvgTuner(0.8), massMatrixTuner=api.IdentityMassMatrixTuner())
seeds = np.arange(64) + 1000
m = api.CudaModel(rir, [], device=0)
src = m.emit_source(cfg)
tr = m.sample(cfg, seeds=seeds, diagnostics=True)
ref = OracleModel(rir, []).sample(api.lower_config(cfg)[0], seeds=seeds)["samples"]
print("equal:", np.array_equal(tr.chains, ref), "rhat", float(tr.diagnostics[:,0].max()), len(src))
