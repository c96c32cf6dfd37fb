"""cuobjdump -sass grl_b200/lib/libgrl_b200.so | python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt
Mnemonic counts per kernel: the tcgen05 / TMA / cluster instructions that prove the Blackwell path, and the absence of legacy HMMA."""
import collections
import re
import sys

per, tot = collections.Counter(), collections.Counter()
pat = re.compile(r"\b(LDTM[.\w]*|UCGABAR_ARV|UCGABAR_WAIT|UTCATOMSWS[.\w]*|UTCBAR[.\w]*|UTCHMMA[.\w]*|UTMALDG[.\w]*|HMMA[.\w]*)")
cur = None
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"^_ZN3grl\d+", "", m.group(1))
        continue
    m = pat.search(line)
    if m and cur:
        per[(cur, m.group(1))] += 1
        tot[m.group(1).split(".")[0] + (".2CTA" if ".2CTA" in m.group(1) else "")] += 1
print("# SASS evidence, libgrl_b200.so at the end of round 2 (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a): cuobjdump -sass, mnemonic counts per kernel")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA; .2CTA = cta_group::2 copies), LDTM = tcgen05.ld,")
print("# UTCBAR = tcgen05.commit (.2CTA.MULTICAST = multicast to the CTA pair), UCGABAR = barrier.cluster; HMMA (legacy mma.sync) must be absent")
for k in sorted(tot):
    print("TOTAL", k, tot[k])
print("TOTAL HMMA", sum(v for (c, m), v in per.items() if m.startswith("HMMA")))
for (c, m), v in sorted(per.items()):
    print(c, m, v)
