"""Per-source-line instruction and stall-sample shares from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    lines = []
    fname, hdr = "?", None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":
            d = dict(zip(hdr, r))
            try:
                inst = int(d["Instructions Executed"] or 0)
                samples = int(d["# Samples"] or 0)
            except ValueError:
                continue
            if inst or samples:
                lines.append((inst, samples, fname, int(r[0]), r[1].strip()))
    total_i = sum(l[0] for l in lines) or 1
    total_s = sum(l[1] for l in lines) or 1
    print(f"total warp instructions {total_i:.4g}, samples {total_s}")
    for inst, samples, f, no, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * inst / total_i:5.1f}% inst {100 * samples / total_s:5.1f}% smp  {f}:{no}  {src[:110]}")


if __name__ == "__main__":
    main()
