"""Aggregates an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]`
launch list per kernel: launches, total time, share, DRAM bytes per launch.
    python tools/ncu_launch_summary.py gpurun_out/launches.csv [--json out.json]
        [--traffic profiles/r01_traffic.json WORKLOAD "source note"]   # refresh bench.py's per-kernel DRAM-traffic table"""
import collections
import csv
import json
import sys


def load(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value) * scale.get(unit, 1)


def to_ns(value, unit):
    scale = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
    return float(value) * scale.get(unit, 1)


def summarise(path):
    per = collections.OrderedDict()
    launches = collections.defaultdict(dict)
    for r in load(path):
        name = r["Kernel Name"].split("::")[-1].split("(")[0].replace("void ", "")
        launches[r["ID"]]["name"] = name
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            launches[r["ID"]]["ns"] = to_ns(r["Metric Value"], r["Metric Unit"])
        elif m.startswith("dram__bytes"):
            launches[r["ID"]][m] = to_bytes(r["Metric Value"], r["Metric Unit"])
    for L in launches.values():
        d = per.setdefault(L["name"], dict(launches=0, ns=0.0, dram_read=0.0, dram_write=0.0))
        d["launches"] += 1
        d["ns"] += L.get("ns", 0.0)
        d["dram_read"] += L.get("dram__bytes_read.sum", 0.0)
        d["dram_write"] += L.get("dram__bytes_write.sum", 0.0)
    return per


if __name__ == "__main__":
    per = summarise(sys.argv[1])
    total = sum(d["ns"] for d in per.values())
    print("| kernel | launches | total ms | share | avg us | DRAM read GB | DRAM write GB | DRAM GB/s (under ncu) |")
    print("|---|---|---|---|---|---|---|---|")
    out = {}
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        gb = (d["dram_read"] + d["dram_write"]) / 1e9
        print(f"| {k} | {d['launches']} | {d['ns'] / 1e6:.3f} | {d['ns'] / total:.3f} | {d['ns'] / d['launches'] / 1e3:.1f} | "
              f"{d['dram_read'] / 1e9:.2f} | {d['dram_write'] / 1e9:.2f} | {gb / (d['ns'] * 1e-9):.0f} |")
        out[k] = dict(launches=d["launches"], total_ms=d["ns"] / 1e6,
                      dram_bytes_per_launch=(d["dram_read"] + d["dram_write"]) / d["launches"])
    print(f"\ntotal kernel time {total / 1e6:.2f} ms")
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        path, workload, note = sys.argv[i + 1], sys.argv[i + 2], sys.argv[i + 3]
        try:
            table = json.load(open(path))
        except FileNotFoundError:
            table = {}
        table[workload] = {k.split("<")[0]: dict(launches=v["launches"], dram_bytes_per_launch=v["dram_bytes_per_launch"],
                                                  source=note) for k, v in out.items() if v["dram_bytes_per_launch"] > 0}
        with open(path, "w") as f:
            json.dump(table, f, indent=1)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)
