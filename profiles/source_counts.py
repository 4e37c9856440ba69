"""Per-source-line warp instruction counts of one kernel from an ncu report taken with --import-source on.
Every SASS instruction is counted once, for the first source line the report lists it under that is not a CUDA header (with inlining the
report repeats an instruction under every line of its inline stack).
usage: python profiles/source_counts.py gpurun_out/x.ncu-rep <cell visits> [min_per_visit] > profiles/xxx_source_counts.txt"""
import csv
import io
import subprocess
import sys


def main():
    rep, visits = sys.argv[1], float(sys.argv[2])
    thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], stdout=subprocess.PIPE, text=True).stdout
    fname, hdr, cur, src = "", None, None, {}
    seen = {}  # SASS address -> [warp instructions, thread instructions, source line key]
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0].isdigit() and hdr:
            cur = (fname, int(r[0]))
            src[cur] = r[1].strip()
        elif r[0] == "" and hdr and cur:
            # (source text may hold unescaped quotes and commas: take the counters from the right end of the row)
            i_n, i_t = hdr.index("Instructions Executed") - len(hdr), hdr.index("Thread Instructions Executed") - len(hdr)
            try:
                n, t = float(r[i_n] or 0), float(r[i_t] or 0)
            except ValueError:
                continue
            addr = r[2]
            if addr not in seen:
                seen[addr] = [n, t, cur]
            elif seen[addr][2][0].endswith(".hpp") and not cur[0].endswith(".hpp"):
                seen[addr][2] = cur
    lines, total = {}, 0.0
    for n, t, key in seen.values():
        a = lines.setdefault(key, [0.0, 0.0])
        a[0] += n
        a[1] += t
        total += n
    print(f"# {rep}: warp instructions per source line / {visits:.0f} cell visits; total {total:.4g} = {total / visits:.1f} per visit")
    print("# file                 line  instr/visit  share  thr/inst  source")
    for (f, ln), (n, t) in sorted(lines.items(), key=lambda kv: -kv[1][0]):
        if n / visits >= thresh:
            print(f"{f:22s} {ln:5d} {n / visits:8.2f} {100 * n / total:6.1f}% {t / n:7.1f}   {src[(f, ln)][:110]}")


if __name__ == "__main__":
    main()
