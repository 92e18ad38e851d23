"""Summarise gpurun_out ncu artefacts into profiles/ (text, committed).
usage: python tools/ncu_summary.py <launches.csv> <prof.ncu-rep> <out.md> [first-kernel-substring] [launches-per-step]"""
import collections, csv, re, subprocess, sys

launch_csv, rep, out = sys.argv[1], sys.argv[2], sys.argv[3]
first_kernel = sys.argv[4] if len(sys.argv) > 4 else "rvq"
per_step = int(sys.argv[5]) if len(sys.argv) > 5 else 101
L = []
lines = [l for l in open(launch_csv) if not l.startswith("==")]
recs = []
for row in csv.DictReader(lines):
    val = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    ns = val * 1e3 if unit.startswith("us") else (val if unit.startswith("ns") else val * 1e6)
    recs.append((row["Kernel Name"], ns, row["Grid Size"], row["Block Size"]))
first = next(i for i, r in enumerate(recs) if first_kernel in r[0])
step = recs[first:first + per_step]
tot = sum(r[1] for r in step)
L.append(f"# ncu launch list: one decode step ({len(step)} launches, {tot/1e6:.3f} ms serialised, cold-cache; compare SHARES)\n")
L.append("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
agg = collections.OrderedDict()
for n, ns, g, b in step:
    k = re.sub(r"\(.*", "", n).replace("void ", "")
    agg.setdefault(k, [0.0, 0]); agg[k][0] += ns; agg[k][1] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    L.append(f"| `{k}` | {v[1]} | {v[0]/1e6:.3f} | {100*v[0]/tot:.1f}% |")
L.append("\n## every launch of the step, in order\n\n| # | kernel | grid | block | us |\n|---:|---|---|---|---:|")
for i, (n, ns, g, b) in enumerate(step):
    L.append(f"| {i} | `{re.sub(r'\(.*', '', n).replace('void ', '')[:60]}` | {g} | {b} | {ns/1e3:.1f} |")
if rep != "-":
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size"]
    L.append(f"\n# ncu --set full capture: {rep}\n")
    L.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |\n|---|---|" + "---|" * (len(rows) - 2))
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            L.append(f"| {w} | {units[i]} | " + " | ".join(r[i][:40] for r in rows[2:]) + " |")
open(out, "w").write("\n".join(L) + "\n")
print("wrote", out)
