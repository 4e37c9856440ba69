"""Per-source-line warp instruction counts of one kernel from an ncu report taken with --import-source on.
usage: python profiles/source_counts.py gpurun_out/x.ncu-rep <visits> [min_per_visit] > profiles/xxx_source_counts.txt"""
import csv
import io
import subprocess
import sys


def main():
    rep, visits = sys.argv[1], float(sys.argv[2])
    thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], stdout=subprocess.PIPE, text=True).stdout
    rows, fname, hdr, total = [], "", None, 0.0
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0].isdigit() and hdr:
            # (source text may hold unescaped quotes and commas: take the counters from the right end of the row)
            i_n, i_t = hdr.index("Instructions Executed") - len(hdr), hdr.index("Thread Instructions Executed") - len(hdr)
            try:
                n, t = float(r[i_n] or 0), float(r[i_t] or 0)
            except ValueError:
                continue
            if n > 0:
                rows.append((n, t, fname, int(r[0]), r[1].strip()))
                total += n
    rows.sort(reverse=True)
    print(f"# {rep}: warp instructions per source line / {visits:.0f} cell visits; total {total:.4g} = {total / visits:.1f} per visit")
    print("# file                 line  instr/visit  share  thr/inst  source")
    for n, t, f, ln, src in rows:
        if n / visits >= thresh:
            print(f"{f:22s} {ln:5d} {n / visits:8.2f} {100 * n / total:6.1f}% {t / n:7.1f}   {src[:110]}")


if __name__ == "__main__":
    main()
