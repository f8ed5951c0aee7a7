"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTCHMMA / UTCQMMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA load / store), SYNCS (mbarrier).
    python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mde_biological_vision_systems_b200", "lib", "libmde_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|LDTM|STTM|UTMALDG(?:\.\dD)?|UTMASTG(?:\.\dD)?|UTMAPF|SYNCS|UTCBAR|HMMA|IMMA)\b")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for tok in pat.findall(line):
            counts[cur][tok] += 1
print("# cuobjdump -sass mde_biological_vision_systems_b200/lib/libmde_b200.so, mnemonic counts per kernel (kernels without any omitted)")
for k, c in counts.items():
    if c:
        print(f"{k[:110]:110s} " + "  ".join(f"{n}={v}" for n, v in sorted(c.items())))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# total: " + "  ".join(f"{n}={v}" for n, v in sorted(tot.items())))
