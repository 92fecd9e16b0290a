"""Dev tool: per-kernel histogram of the Blackwell-specific SASS opcodes in libctcvr.so (cuobjdump -sass): which kernels
issue tcgen05 MMAs (UTC*MMA), tensor-memory loads/stores (LDTM/STTM), TMA / bulk copies (UTMALDG, UBLKCP, UTMASTG), and
whether any legacy tensor path (HMMA) is present.   python tools/sass_hist.py > profiles/r2_sass_histogram.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ctc-vr_b200", "libctcvr.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "MUFU.TANH",
        "MUFU.EX2", "MUFU.LG2", "STSM", "LDSM", "SHFL", "ATOMG", "REDG", "RED.", "ELECT", "UCGABAR", "FFMA2", "LDG.E.64.STRONG", "STG.E.64.STRONG"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0]
        hist[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        hist[cur]["_total"] += 1
        for k in keys:
            if op.startswith(k):
                hist[cur][k] += 1
print("kernel".ljust(70), " ".join(k.rjust(9) for k in ["_total"] + keys))
for name, c in hist.items():
    if c["_total"] == 0:
        continue
    print(name[-70:].ljust(70), " ".join(str(c.get(k, 0)).rjust(9) for k in ["_total"] + keys))
