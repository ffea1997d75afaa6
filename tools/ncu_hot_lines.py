"""Top stall-sample SASS lines of a kernel in an .ncu-rep source page CSV (ncu -i rep --page source --csv)."""
import csv
import sys


def main(path, n=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    i_src = hdr.index("Source"); i_s = hdr.index("Warp Stall Sampling (All Samples)"); i_ex = hdr.index("Instructions Executed")
    data = [r for r in rows[2:] if len(r) > i_ex]
    def iv(x):
        try: return int(x)
        except ValueError: return 0
    tot = sum(iv(r[i_s]) for r in data)
    print("total samples", tot, "lines", len(data))
    top = sorted(range(len(data)), key=lambda i: -iv(data[i][i_s]))[:n]
    for i in sorted(top):
        print(i, data[i][i_s], data[i][i_ex], data[i][i_src][:120])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
