"""Prints the key metrics of every kernel in an .ncu-rep (raw page) as markdown: duration, DRAM bytes, L2/L1 throughput,
tensor-pipe activity, issue utilisation, stall ratios."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum"]


def main(rep, title):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    print(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none --import-source on)\n")
    for row in r[2:]:
        d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
        print(f"## {d.get('Kernel Name', '')[:90]}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k]} | {u[k]} |")
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(d[k]) > 0.2:
                        print(f"| {k.replace('smsp__average_warps_issue_stalled_', 'stall: ').replace('_per_issue_active.ratio', '')} | {d[k]} | warps per issue |")
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu key metrics")
