"""Per-SASS-instruction execution counts of the kernels in an .ncu-rep (source page): python tools/ncu_sass_hot.py report.ncu-rep [kernel_index] [top]
Prints the opcode histogram (warp-level instructions executed) and the hottest instructions."""
import collections
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "cols": None}
            kernels.append(cur)
        elif row[0] == "Address":
            cur["cols"] = row
        elif cur is not None and cur["cols"] is not None:
            cur["rows"].append(row)
    k = kernels[which]
    ci, si, st = k["cols"].index("Instructions Executed"), k["cols"].index("Source"), k["cols"].index("# Samples")
    total = sum(int(r[ci]) for r in k["rows"])
    samples = sum(int(r[st]) for r in k["rows"])
    print(f"kernel {which}: {k['name'][:80]}  SASS instructions {len(k['rows'])}  warp-instructions executed {total}  stall samples {samples}")
    hist, shist = collections.Counter(), collections.Counter()
    for r in k["rows"]:
        op = r[si].split()
        op = [o for o in op if not o.startswith("@")][0].split(".")[0]
        hist[op] += int(r[ci])
        shist[op] += int(r[st])
    print("opcode histogram (share of executed warp-instructions | share of stall samples):")
    for op, n in hist.most_common(28):
        print(f"  {op:10s} {100.0 * n / total:6.2f}%   {100.0 * shist[op] / max(samples, 1):6.2f}%")
    print("hottest instructions by stall samples:")
    for r in sorted(k["rows"], key=lambda r: -int(r[st]))[:top]:
        print(f"  {int(r[st]):7d} smp  {int(r[ci]):10d} exec  {r[si].strip()[:100]}")


if __name__ == "__main__":
    main()
