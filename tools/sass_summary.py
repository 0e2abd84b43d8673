#!/usr/bin/env python
"""Per-kernel count of the tensor-core / TMA / TMEM / vector-reduction SASS instructions of the shipped library
(profiles/r02_sass_summary.txt):   python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fots", "pytorch_b200", "lib", "librroi_b200.so")
KEYS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "FFMA2", "FMUL2", "FHFMA", "REDG", "LDGSTS", "ATOMS", "SYNCS", "UBLKCP")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
print("# cuobjdump -sass fots/pytorch_b200/lib/librroi_b200.so -- tensor-core / TMA / TMEM / vector-reduction instructions per kernel (final round-2 build; tools/sass_summary.py)")
print("# mnemonics: UTCHMMA = tcgen05.mma (kind::f16), .2CTA = cta_group::2; UTMALDG / UTMASTG = TMA tensor load / store; LDTM = tcgen05.ld;")
print("# UTCBAR = tcgen05.commit; HMMA = mma.sync (stem, heads, LSTM, GEMM); FFMA2 / FMUL2 = packed fp32; FHFMA = fma.rn.f32.bf16; REDG...F32x4 = red.global.add.v4.f32;")
print("# LDGSTS = cp.async; ATOMS = shared-memory atomics; SYNCS = mbarrier operations.  Kernels without any of these (the fp32 channels-last forward:\n# plain LDG.128 / FFMA / STG.128) are not listed; one line per template family.\n")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
excerpt = None
seen = collections.Counter()
fam = lambda n: re.sub(r"[<(].*$", "", n.replace("(anonymous namespace)::", "").replace("void ", ""))
for n in names:
    seen[fam(n)] += 1
shown = set()
for name, blk in zip(names, blocks):
    ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Za-z0-9_.]*)", blk)
    cnt = collections.Counter()
    for m in ins:
        for k in KEYS:
            if m.startswith(k):
                cnt[m.replace(".E.", ".").replace(".STRONG.GPU", "") if k == "REDG" else ".".join(m.split(".")[:3]) if k in ("UTCHMMA", "UTMALDG", "UTMASTG") else k] += 1
    if not cnt:
        continue
    short = re.sub(r"\(.*$", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
    is_excerpt_kernel = "conv_tc_kernel<1, 256, 6, true" in name
    if fam(name) in shown and not fam(name).endswith("conv_tc_kernel"):     # one representative per template family
        continue
    shown.add(fam(name))
    extra = "" if seen[fam(name)] == 1 or fam(name).endswith("conv_tc_kernel") else "  (+%d more instantiations)" % (seen[fam(name)] - 1)
    print("%-64s %5d instr | %s%s" % (short[:64], len(ins), ", ".join("%s x%d" % kv for kv in sorted(cnt.items())), extra))
    if excerpt is None and "conv_tc_kernel<1, 256, 6, true" in name:
        lines = [l for l in blk.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        i = next(k for k, l in enumerate(lines) if "UTCHMMA" in l)
        excerpt = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in lines[max(0, i - 12):i + 8]]
if excerpt:
    print("\n# excerpt: MMA issue loop of conv_tc_kernel<1, 256, 6, pair> (around the first UTCHMMA.2CTA)")
    print("\n".join(excerpt))
