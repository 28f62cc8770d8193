"""Summarise ncu outputs into profiles/: (1) a launch list CSV (`--metrics gpu__time_duration.sum`) -> per-kernel
count / total / share; (2) a `--set full` report (`ncu -i rep --page raw --csv`) -> the roofline-relevant metrics."""
import csv, io, re, subprocess, sys, collections

def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols = {c: i for i, c in enumerate(rows[hdr])}
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= cols["Metric Value"] or r[cols["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[cols["Kernel Name"]])
        v = float(r[cols["Metric Value"]].replace(",", ""))
        unit = r[cols["Metric Unit"]]
        v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v_us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot/1000:.3f} ms total (ncu per-launch times are cold-cache and serialised: compare shares)\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>10s} {'share':>7s}\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name[:60]:60s} {n:8d} {us:10.1f} {us/tot:7.3f}\n")

def full(rep, out, pattern):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    want = re.compile(pattern)
    with open(out, "w") as f:
        f.write(f"# {rep}: ncu --set full, one column per captured launch\n")
        idx = [i for i, h in enumerate(hdr) if want.search(h) or h in ("Kernel Name", "Grid Size", "Block Size")]
        for i in idx:
            f.write(f"{hdr[i]} [{units[i]}]: " + " | ".join(r[i] for r in rows[2:]) + "\n")

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        pat = sys.argv[4] if len(sys.argv) > 4 else r"gpu__time_duration.sum|dram__bytes_(read|write)\.sum$|dram__throughput.avg.pct|sm__pipe_tensor.*cycles_active.*pct|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|sm__throughput.avg.pct|lts__t_bytes.sum$|smsp__cycles_active.avg|gpc__cycles_elapsed.max|sm__cycles_elapsed.avg$"
        full(sys.argv[2], sys.argv[3], pat)
