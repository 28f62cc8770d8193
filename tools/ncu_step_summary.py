"""Summaries of ncu CSV logs for profiles/ (round 2).

  launches <csv> <out.txt> [label]   per-kernel count / time / share of ONE steady-state step of bench.py: the launches
                                     between the last two L2-flush fills of the timed loop (bench.py writes 256 MiB between
                                     timed iterations), i.e. no build-time or warm-up launches
  traffic  <csv> <out.json> <workload>  DRAM bytes (read + write, --cache-control none: no per-kernel flush) per
                                     conv_umma_kernel launch over the same step
"""
import collections
import csv
import json
import re
import sys


def rows_of(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols = {c: i for i, c in enumerate(rows[hdr])}
    launches = collections.OrderedDict()   # ID -> dict(name, metrics)
    for r in rows[hdr + 1:]:
        if len(r) <= cols["Metric Value"]:
            continue
        lid = int(r[cols["ID"]])
        d = launches.setdefault(lid, {"name": r[cols["Kernel Name"]], "m": {}})
        v = float(r[cols["Metric Value"]].replace(",", "") or 0)
        unit = r[cols["Metric Unit"]]
        name = r[cols["Metric Name"]]
        if name == "gpu__time_duration.sum":
            v = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        elif unit in ("Kbyte", "KB"):
            v *= 1e3
        elif unit in ("Mbyte", "MB"):
            v *= 1e6
        elif unit in ("Gbyte", "GB"):
            v *= 1e9
        d["m"][name] = v
    return list(launches.values())


def is_flush(l):
    return "FillFunctor" in l["name"] or "fill" in l["name"].lower() and "elementwise" in l["name"].lower()


def last_step(ls):
    """Launches between the last two flush fills that have at least 100 launches between them."""
    idx = [i for i, l in enumerate(ls) if is_flush(l)]
    best = None
    for a, b in zip(idx, idx[1:]):
        if b - a > 100:
            best = (a, b)
    if best is None:
        raise SystemExit("no steady-state step found (no pair of L2-flush fills with a step between them)")
    return ls[best[0] + 1:best[1]]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    name = name.replace("pn::", "").replace("(anonymous namespace)::", "")
    return name


def launches(path, out, label=""):
    step = last_step(rows_of(path))
    agg = collections.OrderedDict()
    for l in step:
        a = agg.setdefault(short(l["name"]), [0, 0.0])
        a[0] += 1
        a[1] += l["m"].get("gpu__time_duration.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    conv = sum(a[1] for k, a in agg.items() if "conv_umma_kernel" in k)
    with open(out, "w") as f:
        f.write(f"# {label or path}\n# one steady-state step: {len(step)} launches, {tot / 1000:.3f} ms summed under ncu "
                "(per-launch times are serialised; --clock-control none): compare SHARES with bench.py's roofline.share_of_serialised_launches\n")
        f.write(f"# conv_umma_kernel (all instantiations): {conv / tot:.3f} of the step\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>10s} {'share':>7s}\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name[:72]:72s} {n:8d} {us:10.1f} {us / tot:7.3f}\n")


def traffic(path, out, workload):
    ls = rows_of(path)
    # only conv launches were captured (-k regex:conv_umma): a step = the last N launches, N = launches per step
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 192
    step = ls[-n:]
    rd = sum(l["m"].get("dram__bytes_read.sum", 0.0) for l in step)
    wr = sum(l["m"].get("dram__bytes_write.sum", 0.0) for l in step)
    us = sum(l["m"].get("gpu__time_duration.sum", 0.0) for l in step)
    json.dump({"workload": workload, "conv_launches_per_step": n, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
               "traffic_bytes_per_launch": (rd + wr) / n, "kernel_us_per_step_under_ncu": us,
               "source": f"{path}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none "
                         "-k regex:conv_umma, last step of the run (L2 NOT flushed per kernel: the step's real DRAM traffic)"},
              open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
