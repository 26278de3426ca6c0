"""Per-source-line execution counts (needs -lineinfo + --import-source on): python tools/ncu_lines_hot.py report.ncu-rep [kernel_index] [top]"""
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    kernels, cur, fpath = [], None, ""
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1].split("/")[-1]
        elif row[0] == "Function Name":
            if not kernels or kernels[-1]["name"] != row[1] or kernels[-1].get("closed"):
                kernels.append({"name": row[1], "lines": []})
            cur = kernels[-1]
        elif row[0] == "Line No":
            cols = row
        elif row[0] == "Kernel Name":
            if kernels:
                kernels[-1]["closed"] = True
        elif cur is not None and row[0] not in ("", "-") and len(row) > 8:
            cur["lines"].append((fpath, row[0], row[1], int(row[7] or 0), int(row[6] or 0)))
    # group consecutive blocks per kernel launch: the page repeats per launch; split when the same (file, line) repeats
    k = kernels[which] if which < len(kernels) else kernels[-1]
    lines = k["lines"]
    total = sum(l[3] for l in lines)
    samples = sum(l[4] for l in lines)
    print(f"{k['name'][:90]}: {len(lines)} source lines, warp-instructions {total}, stall samples {samples}")
    for f, no, src, ex, smp in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f"  {100.0 * ex / max(total, 1):6.2f}% exec  {100.0 * smp / max(samples, 1):6.2f}% smp  {f}:{no}  {src.strip()[:110]}")


if __name__ == "__main__":
    main()
