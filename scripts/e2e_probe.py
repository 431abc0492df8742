import sys, time, ctypes as CT
sys.path.insert(0, '.')
import numpy as np, torch
from rainier_b200 import api, abi
C_, I_ = 151552, 100
rir = open('rainier_b200/models/funnel10.rir','rb').read()
model = api.CudaModel(rir, [], device=0)
cfg = api.make_config(iterations=I_, warmupIterations=0, sampler=api.HMCSampler(5), stepSizeTuner=api.StaticStepSize(0.1), massMatrixTuner=api.IdentityMassMatrixTuner(), launchIterations=I_)
c, keep = api.lower_config(cfg)
seeds = np.arange(C_, dtype=np.int64)
out = np.empty((C_, I_, 10))
for k in range(6):
    t = time.perf_counter()
    rc = api.lib().rn_sample(model.h, CT.byref(c), seeds.ctypes.data, C_, out.ctypes.data, None, None)
    print('call %d rc %d %.1f ms' % (k, rc, (time.perf_counter()-t)*1e3), flush=True)
