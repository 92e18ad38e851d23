"""Per-instruction stall summary of an ncu report (source page): usage: ncu_stalls.py rep [top] [lo hi]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r != hdr]
if len(sys.argv) > 5: data = data[:len(data) // int(sys.argv[5])]   # several launches in one report: keep the first
isamp = hdr.index("# Samples"); isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed")
sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", sum(int(d[isamp]) for d in data), "instructions", len(data))
agg = {}
for d in data:
    for c in sc: agg[hdr[c]] = agg.get(hdr[c], 0) + int(d[c])
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4]); idx = range(lo, hi)
else:
    idx = sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top])
for i in idx:
    d = data[i]
    st = sorted([(int(d[c]), hdr[c][6:]) for c in sc], reverse=True)[:2]
    print(i, d[isrc].strip()[:80], d[isamp], d[iex], st)
# barrier-related instructions
print("--- sync instructions")
for i, d in enumerate(data):
    if any(k in d[isrc] for k in ("TRYWAIT", "UTMALDG", "UTCBAR", "UTMASTG", "UTCHMMA")) and int(d[iex]) > 0:
        print(i, d[isrc].strip()[:80], d[isamp], d[iex])
