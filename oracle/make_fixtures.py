"""
Generates the frozen-DAG fixtures the product-side harnesses (bench.py, smoke) load without touching oracle/:
rainier_b200/models/*.rir, built by the Python restatement of the reference's DAG builder (oracle/rainier_py) in
the reference's exact operation order.  Both flavours are written: `<name>.rir` carries the reference's symbolic
gradient outputs (what Compiler.compileTargets would hand over today), `<name>.primal.rir` only the primal
log-density outputs (what the Scala wrapper of INTEGRATION.md sends).  Data-free models only: streamed models need
their columns, which tests/bench synthesise at run time.
    python oracle/make_fixtures.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.rainier_py import configs  # noqa: E402

OUT = os.path.join(ROOT, "rainier_b200", "models")
os.makedirs(OUT, exist_ok=True)

for name, build in (("funnel10", lambda: configs.funnel(10)), ("eight_schools", configs.eight_schools)):
    m = build()
    rir, cols = m.compile(with_gradient=True)
    assert not cols
    open(os.path.join(OUT, name + ".rir"), "wb").write(rir)
    m = build()
    rir, cols = m.compile(with_gradient=False)
    assert not cols
    open(os.path.join(OUT, name + ".primal.rir"), "wb").write(rir)
    print(name, "ok")

# a compiled function (RIR_FLAG_FUNCTION): the 12 derived quantities of eight schools as Generator.prepare would compile
# them for Trace.predict -- __graft_entry__.build() assembles the function-flavour kernel (rn_k_eval) from it
from oracle.rainier_py.compute import compile_function_rir  # noqa: E402

model, mu, tau, thetas, _ = configs.eight_schools_parts()
open(os.path.join(OUT, "eight_schools.derived.fn.rir"), "wb").write(
    compile_function_rir(model.parameters, configs.eight_schools_derived(mu, tau, thetas)))
print("eight_schools.derived.fn ok")

# one small streamed model with its columns (logistic regression, 700 observations x 4 covariates, primal flavour):
# __graft_entry__.build()/smoke() use it to assemble and run the warp-per-chain (TMA-tiled) kernel shape
import numpy as np  # noqa: E402

rir, cols = configs.logreg(700, 4).compile(False)
np.savez_compressed(os.path.join(OUT, "logreg_700x4.primal.npz"), rir=np.frombuffer(rir, dtype=np.uint8), ncols=len(cols),
                    **{"c%d" % i: np.asarray(c, dtype=np.float64) for i, c in enumerate(cols)})
print("logreg_700x4 ok")
