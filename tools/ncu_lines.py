"""Per-CUDA-source-line totals (warp instructions, stall samples) of one kernel from `ncu --page source --print-source cuda,sass --csv`.

    python tools/ncu_lines.py report.ncu-rep [top_n] [per_unit_divisor]
Needs -lineinfo builds and --import-source on captures.  NCU_KERNEL=<substring> picks a kernel section.
"""
import csv, os, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
want = os.environ.get("NCU_KERNEL", "")
starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
sect = next((i for i in starts if want in rows[i][1]), starts[0])
end = next((i for i in starts if i > sect), len(rows))
print("kernel:", rows[sect][1][:100])
hi = next(i for i in range(sect, end) if rows[i] and rows[i][0] == "Line No")
h = rows[hi]
ci, cs = h.index("Instructions Executed"), h.index("# Samples")
lines = []
for r in rows[hi + 1:end]:
    if len(r) > ci and r[0] not in ("", "Line No") and r[0].isdigit():
        try:
            lines.append((int(r[ci] or 0), int(r[cs] or 0), int(r[0]), r[1].strip()))
        except ValueError:
            pass
ti, ts = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"total warp-instructions {ti/1e6:.1f} M ({ti/div:.0f} per unit), samples {ts}")
for n, s, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"  L{ln:<4d} {n/div:9.1f}/unit {100*n/ti:5.1f} %  samples {100*s/max(ts,1):5.1f} %  {src[:100]}")
