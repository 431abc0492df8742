"""Round-2 diagnosis of test_adaptive_dense_streamed_logistic_regression (warp-per-chain, dense mass adaptation, streamed):
where do GPU and oracle part ways -- smooth rounding drift, or a jump at a window end?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import numpy as np
from oracle.rainier_py import configs
from rainier_b200 import abi, api
import parity

model = configs.logreg(700, 4)
rir, cols = model.compile(True)
for backend, name in ((abi.RN_BACKEND_WARP, "warp"), (abi.RN_BACKEND_THREAD, "thread")):
    for mass, mname in ((api.DenseMassMatrixTuner(30, 1.5, 10, 10), "dense"), (api.DiagonalMassMatrixTuner(30, 1.5, 10, 10), "diag"),
                        (api.IdentityMassMatrixTuner(), "identity")):
        cfg = api.make_config(iterations=20, warmupIterations=120, sampler=api.HMCSampler(3), stepSizeTuner=api.DualAvgTuner(0.8),
                              massMatrixTuner=mass, backend=backend)
        r = parity.run_both(rir, cols, cfg, seeds=np.arange(40) + 9)
        gt, rt = r["gpu_trace"], r["ref_trace"]  # [chains][iters][4]: log accept, accept, step size, steps
        acc_bad = gt[:, :, 1] != rt[:, :, 1]
        rel = np.abs(gt[:, :, 0] - rt[:, :, 0]) / np.maximum(np.abs(rt[:, :, 0]), 1e-12)
        srel = np.abs(gt[:, :, 2] - rt[:, :, 2]) / np.maximum(np.abs(rt[:, :, 2]), 1e-300)
        print("==", name, mname, "accept mismatches", int(acc_bad.sum()), "chains hit", int(acc_bad.any(axis=1).sum()))
        its = [0, 9, 10, 20, 39, 40, 41, 60, 84, 85, 86, 100, 109, 110, 119, 120, 139]
        print("   max rel err of log-accept / step size over chains at iteration:")
        for it in its:
            print("     it %3d  la %.3e  ss %.3e" % (it, np.nanmax(rel[:, it]), np.nanmax(srel[:, it])))
        for c in np.argwhere(acc_bad.any(axis=1))[:4, 0]:
            first = int(np.argmax(acc_bad[c]))
            lo = max(0, first - 3)
            print("   chain", c, "first accept mismatch at it", first)
            for it in range(lo, min(first + 2, gt.shape[1])):
                print("      it", it, "gpu", gt[c, it], "ref", rt[c, it])
        print("   mass rel err", parity.rel_err(r["gpu_mass"], r["ref_mass"]))
