"""
oracle/rainier_py/diagnostics.py -- TEST INFRASTRUCTURE ONLY.  Restatement of Trace.diagnostics
(rainier-core/src/main/scala/com/stripe/rainier/core/Trace.scala:49-121), line by line, sequential sums in the
reference's order.  The reference's tests hold no golden vector for it (parity unpinned: the restatement is the
checker for rn_sampler_diagnostics).
"""
import math


def variogram(trace, lag):  # Trace.scala:111-119
    i = lag
    s = 0.0
    while i < len(trace):
        d = trace[i] - trace[i - lag]
        s += d * d  # Math.pow(x, 2) == x*x exactly
        i += 1
    den = float(len(trace) - lag)
    if den == 0.0:
        return math.nan if s == 0.0 else math.copysign(math.inf, s)
    return s / den


def r_hat_and_v(traces, n, m):  # Trace.scala:63-95
    means = [sum_seq(t) / n for t in traces]
    mean_mean = sum_seq(means) / m
    b = (n / (m - 1)) * sum_seq([(x - mean_mean) * (x - mean_mean) for x in means])
    variances = [sum_seq([(a - mu) * (a - mu) for a in t]) / (n - 1) for t, mu in zip(traces, means)]
    w = sum_seq(variances) / m
    v = (n - 1) / n * w + b / n
    return math.sqrt(v / w), v


def sum_seq(xs):  # Scala's .sum: left fold from 0.0
    s = 0.0
    for x in xs:
        s += x
    return s


def autocorrelation(traces, n, m, v):  # Trace.scala:97-109 (tail recursion as a loop)
    lag, acc = 1, 0.0
    while True:
        vt = sum_seq([variogram(t, lag) for t in traces]) / m
        pt = 1.0 - (vt / (2.0 * v))
        if pt > 0.0 and lag < 100:
            lag, acc = lag + 1, acc + pt
        else:
            return acc


def diagnostics(traces):
    """traces: one sequence of draws per chain (Trace.scala:52-61) -> (rHat, effectiveSampleSize)"""
    m = float(len(traces))
    n = float(len(traces[0]))
    r_hat, v = r_hat_and_v(traces, n, m)
    ac = autocorrelation(traces, n, m, v)
    return r_hat, n * m / (1 + (2 * ac))


def trace_diagnostics(chains):
    """chains: array [chain][iteration][parameter] -> list of (rHat, ess) per parameter (Trace.scala:11-21)"""
    n_par = len(chains[0][0])
    return [diagnostics([[float(a[i]) for a in c] for c in chains]) for i in range(n_par)]
