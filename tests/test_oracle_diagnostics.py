"""CPU checks of the Trace.diagnostics restatement (oracle/rainier_py/diagnostics.py; Trace.scala:49-121): closed-form
cases and the reference's termination quirks.  The reference's own tests hold no golden vector for it."""
import math

import numpy as np

from oracle.rainier_py import diagnostics as D


def test_variogram_and_rhat_closed_form():
    t = [0.0, 1.0, 3.0, 6.0]
    assert D.variogram(t, 1) == (1 + 4 + 9) / 3.0
    assert D.variogram(t, 3) == 36.0
    assert math.isnan(D.variogram(t, 4))          # 0.0 / 0 in the reference
    assert D.variogram(t, 5) == 0.0 and math.copysign(1, D.variogram(t, 5)) == -1.0  # empty sum / negative count
    # two chains with equal means: b = 0 -> v = (n-1)/n * w -> rHat = sqrt((n-1)/n)
    a, b = [1.0, 2.0, 3.0, 4.0], [4.0, 3.0, 2.0, 1.0]
    r, v = D.r_hat_and_v([a, b], 4.0, 2.0)
    w = sum((x - 2.5) ** 2 for x in a) / 3.0
    assert abs(v - 0.75 * w) < 1e-15 and abs(r - math.sqrt(0.75)) < 1e-15


def test_iid_chains_look_converged():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(16, 400))
    r, ess = D.diagnostics([list(c) for c in x])
    assert abs(r - 1.0) < 0.02
    assert 0.5 * x.size < ess < 2.0 * x.size


def test_short_traces_terminate_at_lag_n():
    # a strongly autocorrelated short series keeps pt > 0 until lag == n, where 0/0 = NaN stops the recursion
    x = [[float(i) for i in range(5)], [float(i) + 0.1 for i in range(5)]]
    r, ess = D.diagnostics(x)
    assert math.isfinite(r) and math.isfinite(ess)
