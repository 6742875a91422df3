#!/usr/bin/env python
"""Derives the 31-node Patterson rule (level kgauss=5) that the reference's `quad` routine applies
(hydroscatt/f_2l_tmatrix/tmatrix_py.f:1340-1595) by walking its interleaved node/weight table exactly as the routine
does (index logic of lines 1511-1576), and prints the 15 symmetric (node, weight) pairs + the centre weight in the
order the routine evaluates them.  The numbers are Patterson's published nested Gauss-Kronrod-Patterson rule
(T.N.L. Patterson, Math. Comp. 22 (1968) 847); they are emitted into oracle/patterson31.h and
hydroscatt_b200/csrc/hs_patterson31.h.  Run in the build container (needs /root/reference)."""
import re
import sys

src = open("/root/reference/hydroscatt/f_2l_tmatrix/tmatrix_py.f").read().split("\n")
start = [i for i, l in enumerate(src) if "subroutine quad" in l][0]
txt = "\n".join(src[start:start + 160])
nums = []
for blk in re.findall(r"data d\d/(.*?)/", txt, flags=re.S):
    blk = re.sub(r"\n     \S", " ", blk)
    nums += [float(x.replace("d", "e")) for x in re.findall(r"[-+]?\d\.\d+d[-+]\d+", blk)]
assert len(nums) == 381, len(nums)
p = [None] + nums  # 1-based
kgauss = 5
nodes = [None]  # funct index j -> node
i = 0
iold, inew, k = 0, 1, 2
wt = {}
first = True
while True:
    if not first:
        if k == kgauss:
            break
        k += 1
        wt = {}
        for j in range(1, iold + 1):
            i += 1
            wt[j] = p[i]
    first = False
    iold = iold + inew
    for j in range(inew, iold + 1):
        i += 1
        nodes.append(p[i])
        i += 1
        wt[j] = p[i]
    inew = iold + 1
    i += 1
    wc = p[i]
print("// order of evaluation: centre first, then pairs j=1..15 (sum+x, sum-x)")
print("static const double PAT31_X[15] = {")
print(",\n".join("    %.15e" % nodes[j] for j in range(1, 16)))
print("};\nstatic const double PAT31_W[15] = {")
print(",\n".join("    %.15e" % wt[j] for j in range(1, 16)))
print("};\nstatic const double PAT31_WC = %.15e;" % wc)
s = 2 * sum(wt[j] for j in range(1, 16)) + wc
print("// sum of weights = %.15f (should be 2)" % s, file=sys.stderr)
