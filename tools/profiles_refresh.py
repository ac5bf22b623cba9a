"""Turn ncu outputs (run here, no GPU) into the committed summaries under profiles/.
   launches <csv> <out.md> <title-line>      : per-kernel launch count / total time / share from an `--metrics gpu__time_duration.sum --csv` log
   full <ncu-rep> <out.md> <traffic.json>    : per-launch table of a `--set full` capture + dram bytes per launch per kernel"""
import csv, json, re, subprocess, sys, collections


def short(name):
    """siftb200::<unnamed>::gradient_kernel<false>(PyrView) -> gradient_kernel"""
    head = re.sub(r"<[^<>]*>", "", name.split("(")[0])  # drop template arguments and the <unnamed> namespace
    head = re.sub(r"<[^<>]*>", "", head)
    return head.split("::")[-1].split()[-1] if head.strip() else name[:40]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = rows[0]
    c = {n: i for i, n in enumerate(h)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[c["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[c["Metric Value"]].replace(",", ""))
        unit = r[c["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        k = short(r[c["Kernel Name"]])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += us
    side = lambda k: k.startswith(("exact_", "prep_", "match", "rerank", "void"))  # side measurements / torch fills, outside the timed path
    tot = sum(a[1] for k, a in agg.items() if not side(k))
    with open(out, "w") as f:
        f.write(title + "\n\n| kernel | launches | total us | share of the detect+describe path |\n|---|---|---|---|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {us:.1f} | {'(side measurement)' if side(k) else f'{100 * us / tot:.1f} %'} |\n")
    print(open(out).read())


def full(rep, out, traffic_path):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    c = {n: i for i, n in enumerate(h)}

    def val(r, key, scale=1.0):
        if key not in c or r[c[key]] in ("", "n/a"):
            return None
        v = float(r[c[key]].replace(",", ""))
        u = units[c[key]]
        if u in ("Mbyte",): v *= 1e6
        elif u in ("Kbyte",): v *= 1e3
        elif u in ("Gbyte",): v *= 1e9
        elif u in ("ms", "msecond"): v *= 1e3
        elif u in ("ns", "nsecond"): v *= 1e-3
        elif u in ("s", "second"): v *= 1e6
        return v * scale

    cols = [("time us", "gpu__time_duration.sum", 1), ("dram read MB", "dram__bytes_read.sum", 1e-6), ("dram write MB", "dram__bytes_write.sum", 1e-6),
            ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("issue-active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
            ("warps-active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("FMA pipe cycles %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
            ("LSU wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1), ("warp-instr M", "smsp__inst_executed.sum", 1e-6),
            ("regs", "launch__registers_per_thread", 1)]
    traffic = {}
    with open(out, "a") as f:
        f.write("| kernel | grid | " + " | ".join(n for n, _, _ in cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
        for r in data:
            k = short(r[c["Kernel Name"]])
            cells = []
            for n, key, sc in cols:
                v = val(r, key, sc)
                cells.append("-" if v is None else f"{v:.1f}")
            f.write(f"| {k} | {r[c['launch__grid_size']]} | " + " | ".join(cells) + " |\n")
            t = (val(r, "dram__bytes_read.sum") or 0) + (val(r, "dram__bytes_write.sum") or 0)
            traffic[k] = max(traffic.get(k, 0), int(t))  # a kernel launched per octave: its largest launch
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print(open(out).read()); print(traffic)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
