"""Per-kernel SASS evidence of the Blackwell-native path: counts of tcgen05 MMA (UTC*MMA), TMEM load/store (LDTM/STTM), TMA
(UTMALDG/UTMASTG/UBLKCP), legacy tensor (HMMA), packed fp32x2 (FFMA2/FADD2/FMUL2) and MUFU instructions in libvc_b200.so.

    python tools/sass_summary.py [viewcrafter_b200/libvc_b200.so] > profiles/sass_summary.txt
"""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "viewcrafter_b200", "libvc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
classes = [("tcgen05.mma", r"^UTC\w*MMA"), ("tmem ld/st", r"^(LDTM|STTM)"), ("tma load", r"^(UTMALDG|UBLKCP)"), ("tma store", r"^UTMASTG"),
           ("mma.sync", r"^HMMA"), ("fp32x2", r"^(FFMA2|FADD2|FMUL2)"), ("mufu", r"^MUFU"), ("mbarrier", r"^SYNCS"), ("total", r".")]
rows, cur, cnt = [], None, None
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        for name, pat in classes:
            if re.match(pat, op):
                cnt[name] += 1
if cur:
    rows.append((cur, cnt))
print(f"{os.path.basename(lib)}: SASS architectures {archs}; {len(rows)} kernels")
print(f"{'kernel':70s} " + " ".join(f"{n:>11s}" for n, _ in classes))
tot = collections.Counter()
for name, c in sorted(rows, key=lambda r: -r[1]["total"]):
    d = re.sub(r"\(.*", "", demangle(name))[:70]
    print(f"{d:70s} " + " ".join(f"{c[n]:11d}" for n, _ in classes))
    tot.update(c)
print(f"{'ALL':70s} " + " ".join(f"{tot[n]:11d}" for n, _ in classes))
