"""
oracle/make_bench_models.py -- TEST / BENCH TOOLING (not product): builds the frozen DAGs + observation columns of the BASELINE
configurations that stream data (cfg 2s / 3 / 5) with the Python stand-in of the reference's Scala front end
(oracle/rainier_py: Model.observe, TargetGroup, Translator) and writes them to build/models/<name>.npz, where bench.py and
scripts/bench_configs.py load them.  Model CONSTRUCTION is the reference's job on the JVM ("run the unchanged
Translator, serialize to RIR"); without a JVM in this image the restatement is the only producer of those RIRs.
Run by __graft_entry__.build(); idempotent (existing files are kept).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def streamed_primal(model):
    """the model's primal container WITHOUT the reference's inlining step (rn_model_create inlines on the device)"""
    from oracle.rainier_py import compute
    keep = compute.inlinable
    compute.inlinable = lambda real: False
    try:
        return model.compile(False)
    finally:
        compute.inlinable = keep


def main(which=None):
    from oracle.rainier_py import configs
    d = os.path.join(ROOT, "build", "models")
    os.makedirs(d, exist_ok=True)
    builders = {
        "cfg2s": lambda: configs.linreg(10000, covariates=5).compile(True),
        "cfg2s_primal": lambda: streamed_primal(configs.linreg(10000, covariates=5)),  # what the Scala side sends: no inlining, no gradient
        "cfg3_primal": lambda: configs.logreg(100000, 50).compile(False),
        "cfg4": lambda: configs.eight_schools().compile(True),
        "cfg5_primal": lambda: configs.poisson_glm(1000, 1000000).compile(False),
    }
    for name, build in builders.items():
        if which and name not in which:
            continue
        f = os.path.join(d, name + ".npz")
        if os.path.exists(f):
            continue
        rir, cols = build()
        tmp = f + ".tmp.npz"
        np.savez(tmp, rir=np.frombuffer(rir, dtype=np.uint8), ncols=len(cols), **{"c%d" % i: np.asarray(c, dtype=np.float64) for i, c in enumerate(cols)})
        os.replace(tmp, f)
        print("built", f)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
