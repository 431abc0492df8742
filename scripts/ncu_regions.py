#!/usr/bin/env python
"""Where inside one kernel an `ncu --set full` capture spent its stall samples: the SASS listing of the source page cut into
chunks of 1500 instructions (samples, executed instructions, dominant stall reasons, dominant opcodes), then the hottest loop
(instructions executed > 0.3 x the maximum) by opcode and its 25 most-sampled instructions.
Usage: python scripts/ncu_regions.py gpurun_out/x.ncu-rep > profiles/<tag>_regions.txt"""
import collections, csv, io, subprocess, sys

src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in st}
f = lambda r, i: float(r[i] or 0)
tot, totex = sum(f(r, isamp) for r in data), sum(f(r, iex) for r in data)


def opcode(r, parts=1):
    t = r[isrc].split()
    return ".".join((t[1] if t[0].startswith("@") else t[0]).split(".")[:parts])


print("%d SASS instructions, %d stall samples, %.3e warp instructions executed" % (len(data), tot, totex))
print("chunk   %samples %executed | stall reasons (% of all samples)            | opcodes")
for c in range(0, len(data), 1500):
    ch = data[c:c + 1500]
    reasons = collections.Counter({h[6:]: sum(f(r, idx[h]) for r in ch) for h in st})
    ops = collections.Counter(opcode(r) for r in ch)
    print("%6d  %7.1f %9.1f | %-46s | %s" % (c, 100 * sum(f(r, isamp) for r in ch) / tot, 100 * sum(f(r, iex) for r in ch) / totex,
                                             " ".join("%s:%.1f" % (k, 100 * v / tot) for k, v in reasons.most_common(3)),
                                             " ".join("%s:%d" % kv for kv in ops.most_common(5))))
mx = max(f(r, iex) for r in data)
hot = [i for i, r in enumerate(data) if f(r, iex) > 0.3 * mx]
loop = data[hot[0]:hot[-1] + 1]
ls = sum(f(r, isamp) for r in loop)
print("\nhottest loop: instructions %d..%d, %.1f%% of all samples" % (hot[0], hot[-1], 100 * ls / tot))
byop = collections.defaultdict(lambda: [0.0, 0, collections.Counter()])
for r in loop:
    b = byop[opcode(r, 2)]
    b[0] += f(r, isamp)
    b[1] += 1
    for h in st:
        b[2][h[6:]] += f(r, idx[h])
for op, (s, n, c) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:18]:
    print("  %-16s n=%4d %5.1f%% of loop samples | %s" % (op, n, 100 * s / ls, " ".join("%s:%.1f" % (k, 100 * v / ls) for k, v in c.most_common(3))))
print("\nmost-sampled instructions of the loop")
for i in sorted(sorted(range(len(loop)), key=lambda i: -f(loop[i], isamp))[:25]):
    print("  %6d %5.2f%%  %s" % (hot[0] + i, 100 * f(loop[i], isamp) / ls, loop[i][isrc][:100]))
