#!/usr/bin/env python
"""SASS opcode summary of the kernels NVRTC produces for the BASELINE configurations (no device needed: emit + compile for
sm_100a, cuobjdump -sass): full opcode spellings (so load widths, UBLKCP = the TMA bulk copy, SYNCS = mbarrier, DMMA, MUFU...
are visible) per kernel function.  Writes profiles/<tag>_sass_<config>.txt.   Usage: python scripts/sass_summary.py <tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from rainier_b200 import api

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
N_STEPS = 5


def static(eps, iters, **kw):
    return api.make_config(iterations=iters, warmupIterations=0, sampler=api.HMCSampler(N_STEPS), stepSizeTuner=api.StaticStepSize(eps),
                           massMatrixTuner=api.IdentityMassMatrixTuner(), launchIterations=iters, **kw)


def npz(name):
    z = np.load(os.path.join(ROOT, "build", "models", name + ".npz"))
    return z["rir"].tobytes(), [z["c%d" % i] for i in range(int(z["ncols"]))]


cases = {"funnel": (open(os.path.join(ROOT, "rainier_b200", "models", "funnel10.rir"), "rb").read(), [], static(0.1, 100))}
for name, f, cfg in (("cfg3", "cfg3_primal", static(0.01, 2)), ("cfg5", "cfg5_primal", static(0.0005, 2))):
    if os.path.exists(os.path.join(ROOT, "build", "models", f + ".npz")):
        cases[name] = npz(f) + (cfg,)
for name, (rir, cols, cfg) in cases.items():
    m = api.CudaModel(rir, cols, device=-1)
    cub = "/tmp/sass_%s.cubin" % name
    open(cub, "wb").write(m.emit_cubin(cfg))
    m.close()
    res = subprocess.run(["cuobjdump", "-res-usage", cub], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", cub], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    fn = None
    for l in sass.splitlines():
        mm = re.search(r"Function : (\S+)", l)
        if mm:
            fn = mm.group(1)
            per[fn] = collections.Counter()
            continue
        mm = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
        if mm and fn:
            t = re.sub(r"^@!?U?P\d+\s+", "", mm.group(1).strip())
            per[fn][t.split()[0]] += 1
    out = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, name))
    with open(out, "w") as f:
        f.write("# cuobjdump -sass opcode counts (static), NVRTC sm_100a, config %s\n" % name)
        for l in res.splitlines():
            if "Function" in l or "REG:" in l:
                f.write("# " + l.strip() + "\n")
        for fn, c in per.items():
            if fn not in ("rn_k_iter", "rn_k_warmup", "rn_k_init"):
                continue
            f.write("\n[%s] %d instructions\n" % (fn, sum(c.values())))
            key = [k for k in c if re.match(r"(DMMA|UBLKCP|SYNCS|LDG|STG|LDS|STS|LD\b|LD\.|ST\b|ST\.|LDL|STL|DFMA|DADD|DMUL|MUFU|ATOMS|RED|BAR|SHFL|CALL|BRA)", k)]
            for k in sorted(key, key=lambda k: -c[k]):
                f.write("  %-28s %6d\n" % (k, c[k]))
    print("wrote", out)
