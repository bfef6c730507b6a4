"""profiles/gemm_traffic.json from an `ncu --set full` report of tools/gemm_ncu_compare.py (run in the build container, which has
ncu but no GPU: the .ncu-rep comes back from the GPU box in gpurun_out/).
    python tools/ncu_traffic.py gpurun_out/<name>.ncu-rep
Takes the LAST sf:: GEMM launch of the lm_head shape (16384 x 32000 x 4096, the largest per-TTT-step GEMM) in the report."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(rep):
    if rep.endswith(".csv"):          # the raw page exported on the GPU box (`ncu -i X.ncu-rep --page raw --csv > X.csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    raw = raw[raw.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                      "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size")}
    units = rows[1]
    ours = [r for r in rows[2:] if "sf::gemm" in r[col["Kernel Name"]]]
    if not ours:
        raise SystemExit("no sf::gemm launches in the report")
    r = ours[-1]                      # gemm_ncu_compare.py runs the lm_head shape last

    def gb(name):
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[col[name]]]
        return float(r[col[name]]) * scale

    h = hashlib.sha1()
    for name in ("sf_gemm.cuh", "sf_gemm_wide.cuh", "sf_gemm.cu"):
        with open(os.path.join(ROOT, "specforge_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    M, N, K = 16384, 32000, 4096
    out = {"kernel": r[col["Kernel Name"]].split("(")[0] + " — lm_head forward GEMM 16384 x 32000 x 4096 (bf16 out)",
           "dram_read_bytes": gb("dram__bytes_read.sum"), "dram_write_bytes": gb("dram__bytes_write.sum"),
           "algorithmic_bytes": 2.0 * (M * K + N * K + M * N), "duration_ms": float(r[col["gpu__time_duration.sum"]]),
           "tensor_pipe_active_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
           "source": os.path.basename(rep) + " (ncu --set full --clock-control none, tools/gemm_ncu_compare.py)", "src_sha1": h.hexdigest()}
    with open(os.path.join(ROOT, "profiles", "gemm_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1])
