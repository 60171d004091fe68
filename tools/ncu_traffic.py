"""Turn `ncu --set full` captures into the small per-launch table bench.py reads for the
`roofline.traffic` field (dram__bytes_read.sum + dram__bytes_write.sum per launch).

    python tools/ncu_traffic.py profiles/r01_ncu_traffic.json cap1.ncu-rep [cap2.ncu-rep ...]
"""
import csv
import io
import json
import subprocess
import sys

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}          # durations -> ms


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield {h: (r[i], units[i]) for i, h in enumerate(hdr) if i < len(r)}


def val(cell):
    v, u = cell
    return float(v.replace(",", "")) * UNITS.get(u, 1.0)


def main():
    dst, reps = sys.argv[1], sys.argv[2:]
    table = []
    for rep in reps:
        for r in rows_of(rep):
            name = r["Kernel Name"][0]
            table.append({
                "capture": rep.split("/")[-1], "kernel": name,
                "grid": r.get("launch__grid_size", ("", ""))[0],
                "duration_ms_under_ncu": val(r["gpu__time_duration.sum"]),
                "dram_read_bytes": val(r["dram__bytes_read.sum"]), "dram_write_bytes": val(r["dram__bytes_write.sum"]),
                "dram_bytes": val(r["dram__bytes_read.sum"]) + val(r["dram__bytes_write.sum"]),
                "tensor_pipe_active_pct": float(r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", ("0", ""))[0] or 0),
                "dram_throughput_pct": float(r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("0", ""))[0] or 0),
            })
    json.dump({"how": "ncu --set full --clock-control none, one launch each; bytes are per launch", "launches": table}, open(dst, "w"), indent=1)
    for t in table:
        print(f'{t["kernel"][:70]:70s} {t["dram_bytes"] / 1e6:10.1f} MB  {t["duration_ms_under_ncu"]:8.3f} ms  tensor {t["tensor_pipe_active_pct"]:5.1f}%')


if __name__ == "__main__":
    main()
