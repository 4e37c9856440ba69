"""Launch list (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) -> per-kernel table.
usage: python profiles/step_breakdown.py gpurun_out/launches.csv <steps in the run> > profiles/xxx_step_breakdown.txt"""
import csv
import re
import sys


def main():
    path, steps = sys.argv[1], float(sys.argv[2])
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    per = {}  # launch id -> [name, ns, bytes]
    for r in rows:
        e = per.setdefault(int(r[0]), [r[4], 0.0, 0.0])
        v = float(r[14].replace(",", ""))
        if r[12] == "gpu__time_duration.sum":
            e[1] = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[13], 1.0)
        else:
            e[2] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[13], 1.0)
    agg, order = {}, []
    for _, (name, ns, b) in sorted(per.items()):
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)[:52]
        if short not in agg:
            agg[short] = [0, 0.0, 0.0]
            order.append(short)
        a = agg[short]
        a[0] += 1
        a[1] += ns
        a[2] += b
    print("# kernel                                               launches   total_us   us/launch   us/step   dram MB/launch   GB/s")
    for k in order:
        n, ns, b = agg[k]
        print(f"  {k:52s} {n:5d} {ns / 1e3:11.1f} {ns / 1e3 / n:10.1f} {ns / 1e3 / steps:9.1f} {b / 1e6 / n:14.1f} {b / max(ns, 1):8.0f}")


if __name__ == "__main__":
    main()
