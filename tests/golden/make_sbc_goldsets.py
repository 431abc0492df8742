"""
Extracts the reference's own golden vectors for the HMC path into tests/golden/sbc_goldsets.json.

Source: /root/reference/rainier-test/src/main/scala/com/stripe/rainier/core/SBCModel.scala:46-267 (the `goldset`
lists; compared at 1e-10 relative by rainier-test/src/test/scala/com/stripe/rainier/core/SBCTest.scala:7-15).
Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_sbc_goldsets.py
"""
import json
import os
import re

SRC = "/root/reference/rainier-test/src/main/scala/com/stripe/rainier/core/SBCModel.scala"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sbc_goldsets.json")

text = open(SRC).read()
out = {}
for m in re.finditer(r"object (SBC\w+) extends SBCModel\[\w+\] \{(.*?)\n\}", text, re.S):
    name, body = m.group(1), m.group(2)
    g = re.search(r"def goldset =\s*List\((.*?)\)", body, re.S)
    line = text[: m.start()].count("\n") + 1
    vals = [float(v) for v in re.findall(r"-?\d+\.\d+(?:[eE]-?\d+)?", g.group(1))]
    out[name] = {"source_line": line, "goldset": vals}
json.dump({"seed": 1528673302081, "warmup": 10000, "synthetic_samples": 1000, "models": out}, open(OUT, "w"), indent=1)
print({k: len(v["goldset"]) for k, v in out.items()})
