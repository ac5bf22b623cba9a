"""Dynamic instruction mix + hottest SASS lines of one kernel from `ncu --page source --csv` (run here)."""
import csv, subprocess, sys, collections, re

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sys.argv[3:], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one section per captured kernel ("Kernel Name" row, then a header row starting with "Address"); NCU_KERNEL picks one by substring
import os
want = os.environ.get("NCU_KERNEL", "")
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sect = next((i for i in starts if want in rows[i][1]), starts[0])
end = next((i for i in starts if i > sect), len(rows))
rows = rows[sect:end]
print("kernel:", rows[0][1][:90])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
c = {n: i for i, n in enumerate(h)}
mix = collections.Counter(); samples = collections.Counter(); wave = collections.Counter(); ideal = collections.Counter()
tot = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(h) or not r[c["Source"]]:
        continue
    src = r[c["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else src.split()[0]
    n = int(r[c["Instructions Executed"]] or 0)
    s = int(r[c["# Samples"]] or 0)
    mix[op] += n; samples[op] += s; tot += n
    if "L1 Wavefronts Shared" in c:  # absent for kernels without shared-memory traffic
        wave[op] += int(r[c["L1 Wavefronts Shared"]] or 0); ideal[op] += int(r[c["L1 Wavefronts Shared Ideal"]] or 0)
    lines.append((s, n, src))
ts = sum(samples.values())
print(f"total warp-instructions {tot/1e6:.1f} M, samples {ts}")
for op, n in mix.most_common(18):
    extra = f"  smem wavefronts {wave[op]/1e6:.1f} M (ideal {ideal[op]/1e6:.1f} M)" if wave[op] else ""
    print(f"  {op:10s} {n/1e6:9.2f} M  {100*n/tot:5.1f} %   samples {100*samples[op]/max(ts,1):5.1f} %{extra}")
for s, n, src in sorted(lines, reverse=True)[:top]:
    print(f"  {s:6d} samples  {n/1e6:7.2f} M  {src[:110]}")
