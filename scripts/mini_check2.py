"""torch-free GPU sanity of the rows added last (a few seconds): rn_function_eval and rn_optimize (thread and warp shape)
against the oracle on the committed fixtures / a tiny streamed model.  Prints one line per check."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
t0 = time.time()
from oracle.rainier_py.binding import OracleFunction, OracleModel  # noqa: E402
from oracle.rainier_py.optimizer import lbfgs  # noqa: E402
from rainier_b200 import abi, api  # noqa: E402

frir = open("rainier_b200/models/eight_schools.derived.fn.rir", "rb").read()
x = np.random.default_rng(0).normal(size=(1000, 10))
f = api.CudaFunction(frir)
print("function equal:", bool(np.array_equal(f(x), OracleFunction(frir)(x), equal_nan=True)), "%.1fs" % (time.time() - t0), flush=True)
rir = open("rainier_b200/models/eight_schools.rir", "rb").read()
cm, om = api.CudaModel(rir, []), OracleModel(rir, [])
x0 = np.random.default_rng(1).normal(size=(64, 10)) * 0.7
x0[0] = 0
got = cm.optimize(x0, max_evals=300, backend=abi.RN_BACKEND_THREAD)
ref = [lbfgs(om.density_batch, 10, x0=v, max_evals=300) for v in x0[:8]]
print("lbfgs thread equal:", all(np.array_equal(got["x"][c], np.array(r["x"]), equal_nan=True) and got["evals"][c] == r["evals"]
                                 for c, r in enumerate(ref)), "%.1fs" % (time.time() - t0), flush=True)
pred, tr = cm.sample_predict(f, api.SamplerConfig(iterations=10, warmupIterations=60), seeds=np.arange(96) + 1)
full = cm.sample(api.SamplerConfig(iterations=10, warmupIterations=60), seeds=np.arange(96) + 1)
print("sample_predict equal:", bool(np.array_equal(pred.reshape(-1, 12), f(full.chains.reshape(-1, 10)), equal_nan=True)),
      "%.1fs" % (time.time() - t0), flush=True)
z = np.load("rainier_b200/models/logreg_700x4.primal.npz")
cols = [z["c%d" % i] for i in range(int(z["ncols"]))]
w = api.CudaModel(z["rir"].tobytes(), cols).optimize(np.zeros((4, 4)), eps=1e-5, max_evals=300, backend=abi.RN_BACKEND_WARP)
print("lbfgs warp:", w["info"].tolist(), w["evals"].tolist(), float(np.ptp(w["x"], axis=0).max()), "%.1fs" % (time.time() - t0), flush=True)
