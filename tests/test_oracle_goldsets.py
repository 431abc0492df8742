"""
Pins the CPU oracle (oracle/rainier_oracle.cpp driven through oracle/rainier_py) to the reference's OWN golden
vectors: rainier-test/src/main/scala/com/stripe/rainier/core/SBCModel.scala:46-267, compared exactly as
rainier-test/src/test/scala/com/stripe/rainier/core/SBCTest.scala:7-15 does (relative error < 1e-10), for the
enabled list of SBCTest.scala:20-34.  Each case runs: ScalaRNG(1528673302081) -> synthesize 1000 observations
-> Model.observe -> HMCSampler(1)/DualAvgTuner(0.8)/Identity, 10000 warmup + len(goldset) iterations, 1 chain on
the same RNG stream -> predict.
"""
import json
import os

import pytest

from oracle.rainier_py import sbc_models

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sbc_goldsets.json")))


@pytest.mark.parametrize("name", sbc_models.ENABLED)
def test_goldset(name):
    gold = GOLD["models"][name]["goldset"]
    out = sbc_models.run(name, len(gold), seed=GOLD["seed"], synthetic_samples=GOLD["synthetic_samples"],
                         warmup=GOLD["warmup"])
    assert len(out) == len(gold)
    for a, b in zip(out, gold):
        assert abs((a - b) / b) < 1e-10
