"""Counts the Blackwell-specific SASS mnemonics per kernel family in libunet3d_b200.so (cuobjdump -sass) -> markdown."""
import collections
import re
import subprocess
import sys

MNE = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "UTCBAR", "HMMA"]


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fam = None
    cnt = collections.defaultdict(lambda: collections.Counter())
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            k = re.search(r"(conv_\w+?_kernel|conv_igemm_kernel|loss_\w+_kernel|head_fwd_kernel|norm_act_\w+_kernel|channel_reduce_kernel|k_warp)", name)
            fam = k.group(1) if k else None
            continue
        if fam is None:
            continue
        for mn in MNE:
            if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
                cnt[fam][mn] += 1
    print("# SASS evidence (cuobjdump -sass unet-studio_b200/libunet3d_b200.so, sm_100a)\n")
    print("Instruction mnemonics per kernel family, summed over template instantiations.  `UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` = "
          "tcgen05.ld, `UTMALDG` = cp.async.bulk.tensor (TMA tensor-map load), `UBLKCP` = cp.async.bulk (1-D bulk copy on an mbarrier), "
          "`LDGSTS` = cp.async 16 B with zero fill, `SYNCS` = mbarrier ops, `UTCBAR` = tcgen05.commit.  `HMMA` (legacy mma.sync path) "
          "does not occur.\n")
    print("| kernel | " + " | ".join(MNE) + " |\n|---|" + "---:|" * len(MNE))
    for f in sorted(cnt):
        if any(cnt[f][m] for m in MNE[:4]):
            print(f"| `{f}` | " + " | ".join(str(cnt[f][m]) for m in MNE) + " |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "unet-studio_b200/libunet3d_b200.so")
