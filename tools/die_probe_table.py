"""Join the ncu log of tools/microbench/die_probe (launch order: for s in SMs: touch(0), touch(s)) with the die map that
evo_b200/csrc/die_map.cu dumped (EVO_B200_GEMM_DIE_DUMP): DRAM bytes of the second read per SM, grouped by measured die.

    python tools/die_probe_table.py gpurun_out/r02_die_probe.csv gpurun_out/r02_die_map.txt"""
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    while rows[0][0] != "ID":
        rows.pop(0)
    ix = {h: i for i, h in enumerate(rows[0])}
    per = {}
    for r in rows[1:]:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[r[ix["Metric Unit"]]]
        per.setdefault(int(r[ix["ID"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale
    ids = sorted(per)
    first = [per[i] for i in ids[0::2]]
    second = [per[i] for i in ids[1::2]]
    die = {}
    verdict = "no dump"
    if len(sys.argv) > 2:
        lines = open(sys.argv[2]).read().splitlines()
        verdict = lines[0]
        die = {int(l.split()[0]): int(l.split()[1]) for l in lines if l and not l.startswith("#")}
    print("die map:", verdict, "| SMs on die 0 / 1:", sum(1 for d in die.values() if d == 0), "/", sum(1 for d in die.values() if d == 1))
    print("first read (SM 0, cold L2): mean %.1f MB" % (sum(m["dram__bytes_read.sum"] for m in first) / len(first) / 1e6))
    for d in (0, 1, -1):
        sel = [s for s in range(len(second)) if die.get(s, -1) == d]
        if not sel:
            continue
        b = [second[s]["dram__bytes_read.sum"] / 1e6 for s in sel]
        t = [second[s]["gpu__time_duration.sum"] for s in sel]
        print("second read from SMs of die %2d (%3d SMs): DRAM read mean %.2f MB, min %.2f, max %.2f; duration mean %.0f us" % (d, len(sel), sum(b) / len(b), min(b), max(b), sum(t) / len(t)))
    if "-v" in sys.argv:
        for s, m in enumerate(second):
            print(s, die.get(s, -1), "%.2f MB" % (m["dram__bytes_read.sum"] / 1e6), "%.0f us" % m["gpu__time_duration.sum"])


if __name__ == "__main__":
    main()
