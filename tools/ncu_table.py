"""Markdown table of the key metrics of every launch in an ncu report. usage: ncu_table.py rep [labels,comma,separated]"""
import csv, subprocess, sys
rep = sys.argv[1]
labels = sys.argv[2].split(",") if len(sys.argv) > 2 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts %"),
        ("launch__registers_per_thread", "regs"), ("launch__block_size", "threads")]
idx = [(hdr.index(m), n) for m, n in want if m in hdr]
print("| launch | " + " | ".join(f"{n} ({units[i]})" if units[i] else n for i, n in idx) + " |")
print("|---|" + "---|" * len(idx))
for k, r in enumerate(rows[2:]):
    lab = labels[k] if k < len(labels) else r[hdr.index("Kernel Name")].split("::")[-1][:28]
    vals = []
    for i, n in idx:
        try: vals.append(f"{float(r[i].replace(',', '')):.3f}".rstrip("0").rstrip("."))
        except ValueError: vals.append(r[i])
    print(f"| {lab} | " + " | ".join(vals) + " |")
