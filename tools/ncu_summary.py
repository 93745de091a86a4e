"""Summarise an `ncu --set full` report into profiles/: one CSV row per profiled launch with the metrics the
roofline discussion uses, and profiles/ncu_traffic.json (kernel -> DRAM bytes per launch) for bench.py.

    python tools/ncu_summary.py gpurun_out/prof_train_r1.ncu-rep profiles/r1_train_ncu.csv [--traffic replay]
"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("ar::", "")
    return n.split("<")[0].replace("_kernel", "")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    mode = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    traffic = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s [%s]" % (m, units[i]) for m, i in cols])
        for r in data:
            w.writerow([short(r[ik])] + [r[i] for _, i in cols])
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")]) * UNIT_SCALE[units[hdr.index("dram__bytes_read.sum")]]
                wr = float(r[hdr.index("dram__bytes_write.sum")]) * UNIT_SCALE[units[hdr.index("dram__bytes_write.sum")]]
                traffic.setdefault(short(r[ik]), []).append(rd + wr)
            except (ValueError, KeyError):
                pass
    print("wrote", out, "(%d launches)" % len(data))
    if mode:
        tp = os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_traffic.json")
        d = json.load(open(tp)) if os.path.exists(tp) else {}
        d[mode] = {k: sum(v) / len(v) for k, v in traffic.items()}
        d.setdefault("_source", {})[mode] = os.path.basename(rep)
        json.dump(d, open(tp, "w"), indent=1, sort_keys=True)
        print("updated", tp)


if __name__ == "__main__":
    main()
