"""Round-2 diagnosis of the rn_k_eval bit mismatches: dump the first mismatching rows/outputs."""
import struct
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import numpy as np
from oracle.rainier_py.binding import OracleFunction
from oracle.rainier_py.compute import compile_function_rir
from rainier_b200 import api
from test_function_host import _derived, schools


def hx(v):
    return struct.pack(">d", float(v)).hex()


model, mu, tau, thetas, _ = schools()
reals = _derived(mu, tau, thetas)
rir = compile_function_rir(model.parameters, reals)
for count in (128, 129, 4097, 300000):
    rows = np.random.default_rng(count).normal(size=(count, 10)) * 1.3
    rows[0, 0], rows[1, 1], rows[2, 2] = np.nan, np.inf, -np.inf
    f = api.CudaFunction(rir)
    got = f(rows)
    ref = OracleFunction(rir)(rows)
    bad = ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
    idx = np.argwhere(bad)
    print("count", count, "mismatches", len(idx), "outputs hit", sorted(set(idx[:, 1].tolist())))
    for r, j in idx[:12]:
        print("  row", r, "out", j, "got", got[r, j], hx(got[r, j]), "ref", ref[r, j], hx(ref[r, j]), "in", [hx(v) for v in rows[r]])
