#!/usr/bin/env python
"""Static SASS of rn_k_iter for cfg3 / cfg5 (no device): instruction count, control flow, registers, per environment setting.
Usage: RN_ROW_LIBM=0|1 python scripts/r2b/sass_cfg.py cfg3 cfg5"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from rainier_b200 import api
for name in sys.argv[1:]:
    z = np.load(os.path.join(ROOT, "build", "models", name + "_primal.npz"))
    cols = [z["c%d" % i] for i in range(int(z["ncols"]))]
    m = api.CudaModel(z["rir"].tobytes(), cols, device=-1)
    cfg = api.make_config(iterations=2, warmupIterations=0, sampler=api.HMCSampler(5), stepSizeTuner=api.StaticStepSize(0.01),
                          massMatrixTuner=api.IdentityMassMatrixTuner(), launchIterations=2)
    cub = "/tmp/sass_%s.cubin" % name
    open(cub, "wb").write(m.emit_cubin(cfg))
    m.close()
    res = subprocess.run(["cuobjdump", "-res-usage", cub], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "rn_k_iter", cub], capture_output=True, text=True).stdout
    ops = [re.sub(r"^@!?U?P\d+\s+", "", mm.group(1).strip()).split()[0].split(".")[0]
           for mm in (re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l) for l in sass.splitlines()) if mm]
    c = collections.Counter(ops)
    regs = re.search(r"Function rn_k_iter:\s*\n?\s*REG:(\d+) STACK:(\d+)", res)
    print(name, "RN_ROW_LIBM=%s" % os.environ.get("RN_ROW_LIBM", "1"), "regs/stack", regs.groups() if regs else None, "total", len(ops),
          "fp64", sum(c[k] for k in ("DADD", "DMUL", "DFMA", "DSETP", "DMMA")), "BRA", c["BRA"], "BSSY", c["BSSY"], "CALL", c["CALL"],
          "MUFU", c["MUFU"], "UMOV", c["UMOV"], "LDCU", c["LDCU"])
