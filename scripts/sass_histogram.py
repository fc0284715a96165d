"""Per-kernel histogram of the Blackwell-specific SASS opcodes of the built library (cuobjdump -sass), the evidence that
the hot kernels are tcgen05 / TMEM / TMA code:  UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld /
st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce-add), SYNCS (mbarrier), plus HMMA / FFMA counts for contrast.
    python scripts/sass_histogram.py [lib.so] > profiles/rNN_sass_histogram.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "hoisdf_b200/libhoisdf_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMALDG.MULTICAST", "UTMASTG", "UTMAREDG", "SYNCS", "HMMA", "FFMA", "total"]
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern.replace("(anonymous namespace)::", "")).replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        h = hist[kern]
        h["total"] += 1
        base = op.split(".")[0]
        if base in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "HMMA", "FFMA"):
            h[base] += 1
        if base == "UTCHMMA" and ".2CTA" in op: h["UTCHMMA.2CTA"] += 1
        if base == "UTMALDG" and "MULTICAST" in op: h["UTMALDG.MULTICAST"] += 1
print("SASS opcode histogram of %s (sm_100a); one row per kernel" % lib)
print("%-64s " % "kernel" + " ".join("%9s" % o[-9:] for o in OPS))
tot = collections.Counter()
for k, h in hist.items():
    tot.update(h)
    print("%-64s " % k[:64] + " ".join("%9d" % h[o] for o in OPS))
print("%-64s " % "ALL" + " ".join("%9d" % tot[o] for o in OPS))
