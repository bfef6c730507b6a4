"""Per-kernel summary of an `ncu --set full` capture: one CSV row per launch with the handful of metrics the roofline table in
profiles/README.md quotes, plus a per-kernel-name aggregate on stderr.

    python tools/ncu_summary.py <report.ncu-rep | raw.csv> [out.csv]

Accepts either the .ncu-rep (read with `ncu -i ... --page raw --csv`, works without a GPU) or that raw CSV itself — on the GPU box
the raw page is written and the multi-hundred-MB report deleted, so only the CSV travels back."""
import csv
import io
import subprocess
import sys

KEEP = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "xu_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp_inst"),
        ("sm__cycles_elapsed.avg.per_second", "sm_hz"), ("lts__t_sectors_srcunit_tex.sum", "l2_sectors_from_sm")]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6,
         "nsecond": 1e-9, "second": 1.0, "Ghz": 1e9, "Mhz": 1e6, "hz": 1.0, "cycle/nsecond": 1e9, "cycle/usecond": 1e6, "cycle/second": 1.0}


def rows_of(path):
    if path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    raw = raw[raw.index('"ID"'):]                       # drop ==PROF== banner lines
    return list(csv.reader(io.StringIO(raw)))


def short(name):
    name = name.split("(")[0].replace("void ", "").replace("sf::", "")
    return name.strip()


def main():
    rows = rows_of(sys.argv[1])
    hdr, units, data = rows[0], rows[1], rows[2:]
    kcol = hdr.index("Kernel Name")
    cols = [(hdr.index(m), tag, units[hdr.index(m)]) for m, tag in KEEP if m in hdr]
    out = [["id", "kernel"] + [tag + ("_s" if tag == "time" else "_bytes" if tag.startswith("dram_") and tag != "dram_pct" else "") for _, tag, _ in cols]
           + ["dram_GBps"]]
    agg = {}
    for i, r in enumerate(data):
        vals = {}
        for c, tag, u in cols:
            try:
                v = float(r[c].replace(",", ""))
            except ValueError:
                v = float("nan")
            vals[tag] = v * SCALE.get(u, 1.0)
        bw = (vals.get("dram_read", 0) + vals.get("dram_write", 0)) / vals["time"] / 1e9 if vals.get("time") else float("nan")
        name = short(r[kcol])
        out.append([i, name] + [f"{vals[tag]:.6g}" for _, tag, _ in cols] + [f"{bw:.1f}"])
        a = agg.setdefault(name, {"n": 0, "time": 0.0, "bytes": 0.0, "tensor": 0.0, "dram_pct": 0.0, "regs": vals.get("regs", 0)})
        a["n"] += 1; a["time"] += vals.get("time", 0); a["bytes"] += vals.get("dram_read", 0) + vals.get("dram_write", 0)
        a["tensor"] += vals.get("tensor_pct", 0); a["dram_pct"] += vals.get("dram_pct", 0)
    w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
    w.writerows(out)
    print(f"{'kernel':58s} {'n':>4s} {'mean ms':>9s} {'GB/launch':>10s} {'GB/s':>8s} {'dram %':>7s} {'tensor %':>9s} {'regs':>5s}", file=sys.stderr)
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["time"]):
        n = a["n"]
        print(f"{name[:58]:58s} {n:4d} {a['time'] / n * 1e3:9.4f} {a['bytes'] / n / 1e9:10.4f} {a['bytes'] / a['time'] / 1e9 if a['time'] else 0:8.0f} "
              f"{a['dram_pct'] / n:7.1f} {a['tensor'] / n:9.1f} {int(a['regs']):5d}", file=sys.stderr)


if __name__ == "__main__":
    main()
