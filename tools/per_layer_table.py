"""Per-layer tensor-pipe table: joins the library's launch trace (U3D_TRACE_LAUNCHES, one line + one marker kernel per tensor-kernel
launch site) with an ncu launch list of the same run taken with
   --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
(U3D_ONE_STREAM=1 U3D_NO_GRAPH=1 so that the launch order is the trace order).  Prints markdown."""
import collections
import csv
import sys


def main(trace_path, csv_path, title, step_index=1):
    trace = [l.rstrip("\n").split("\t") for l in open(trace_path)]
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        k = r["ID"]
        d = per.setdefault(k, {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    kern = list(per.values())
    # split at markers
    groups, cur = [], None
    for k in kern:
        if "trace_marker" in k["name"]:
            cur = []
            groups.append(cur)
        elif cur is not None and any(t in k["name"] for t in ("conv_band", "conv_tma", "conv_s2", "conv_wgrad", "conv_igemm", "conv_zband", "splitk", "wgrad_band_sum")):
            cur.append(k)
    n = min(len(groups), len(trace))
    # one training step = the entries between two "encode0.0.weight fwd" lines
    starts = [i for i in range(n) if trace[i][0] == "encode0.0.weight" and trace[i][1] == "fwd"]
    lo = starts[step_index] if len(starts) > step_index else starts[0]
    hi = starts[step_index + 1] if len(starts) > step_index + 1 else n
    print(f"# {title}\n")
    print("one training micro-batch, launch order; time and tensor-pipe activity from ncu (serialised, cold cache: compare shares), "
          "FLOPs = algorithmic 2*Cin*Cout*k^3*V_out\n")
    print("| layer | pass | kernel | shape | launches | us | TFLOP/s | tensor pipe active % |\n|---|---|---|---|---:|---:|---:|---:|")
    tot_t = tot_f = wsum = 0.0
    fam = collections.OrderedDict()
    for i in range(lo, hi):
        name, pas, family, nprob, flops, shape = trace[i]
        g = groups[i]
        t = sum(k.get("gpu__time_duration.sum", 0.0) for k in g) / 1e3
        main_k = [k for k in g if "splitk" not in k["name"] and "_sum_kernel" not in k["name"]]
        tp = (sum(k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * k.get("gpu__time_duration.sum", 0.0) for k in main_k) /
              max(sum(k.get("gpu__time_duration.sum", 0.0) for k in main_k), 1e-9))
        fl = float(flops)
        print(f"| {name.replace('.weight', '')} | {pas} | {family} | {shape} | {len(g)} | {t:.1f} | {fl / 1e12 / (t / 1e6):.0f} | {tp:.1f} |")
        tot_t += t; tot_f += fl; wsum += tp * fl
        a = fam.setdefault(family, [0.0, 0.0, 0.0])
        a[0] += t; a[1] += fl; a[2] += tp * fl
    print(f"\nFLOP-weighted tensor-pipe activity over the step's tensor kernels: **{wsum / tot_f:.1f} %**; {tot_f / 1e9:.0f} GFLOP in {tot_t / 1e3:.2f} ms "
          f"of tensor-kernel time = {tot_f / 1e12 / (tot_t / 1e6):.0f} TFLOP/s\n")
    print("| kernel family | us | GFLOP | TFLOP/s | FLOP-weighted tensor pipe % |\n|---|---:|---:|---:|---:|")
    for k, a in fam.items():
        print(f"| {k} | {a[0]:.0f} | {a[1] / 1e9:.0f} | {a[1] / 1e12 / (a[0] / 1e6):.0f} | {a[2] / a[1]:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "per-layer tensor-pipe table")
