"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel summary (markdown) of one complete
step in the capture (from one pack_act_kernel launch = start of a forward to the next)."""
import collections
import csv
import sys


def main(path, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    idx = [i for i, r in enumerate(rows) if "pack_act" in r["Kernel Name"]]
    # a complete step = from one pack_act (start of a forward) to the next; fall back to the tail of the capture
    seg = rows[idx[0]:idx[1]] if len(idx) >= 2 else (rows[idx[-1]:] if idx else rows)
    seg = [r for r in seg if "at::" not in r["Kernel Name"]]
    agg = collections.OrderedDict()
    for r in seg:
        n = r["Kernel Name"].split("(")[0].replace("void ", "").replace("u3d::<unnamed>::", "")
        t = float(r["Metric Value"].replace(",", "")) / 1e3
        a = agg.setdefault(n, [0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] = max(a[2], t)
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n\nsource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache serialised launches: compare shares)\n")
    print(f"launches: {len(seg)}, sum of kernel time: {tot/1e3:.2f} ms\n")
    print("| kernel | launches | total us | share | longest us |\n|---|---:|---:|---:|---:|")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {v[0]} | {v[1]:.0f} | {100*v[1]/tot:.1f}% | {v[2]:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
