#!/usr/bin/env python
"""Copy one evidence pass (gpurun_out/<tag>_*: bench line, launch list, per-launch DRAM list, ncu --set full logs, per-layer
table, tuned-tile table, pytest log) into profiles/r2_* and print the family shares of the launch list.
    python tools/refresh_profiles.py r2ac"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = (("conv_tc", "conv fwd/dgrad (tcgen05)"), ("wgrad_tc", "conv wgrad (tcgen05)"), ("splitk", "split-K reduce"), ("sum_parts", "split-K reduce"),
            ("gn_", "norm family"), ("warp_forward", "warp forward"), ("warp_backward", "warp backward"), ("fill4", "warp backward"),
            ("nnloss", "nn_loss"), ("adam", "Adam"), ("pack", "weight pack/transposes"), ("transpose", "weight pack/transposes"),
            ("unpack", "weight pack/transposes"), ("head", "1x1/head helpers"), ("narrow", "narrow convs (SIMT)"), ("nchw", "layout"),
            ("gather_nhwc", "layout"), ("mask_pyramid", "mask pyramid"), ("at::", "torch helper kernels"))


def main():
    tag = sys.argv[1]
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    ext = os.path.join(ROOT, "tools", "ncu_extract.py")
    shutil.copy(os.path.join(go, tag + "_launches.csv"), os.path.join(pr, "r2_launches_step_256x256_n8.csv"))
    shutil.copy(os.path.join(go, tag + "_dram.csv"), os.path.join(pr, "r2_dram_per_launch.csv"))
    shutil.copy(os.path.join(go, tag + "_layers.txt"), os.path.join(pr, "r2_conv_layers.txt"))
    shutil.copy(os.path.join(go, tag + "_tune.txt"), os.path.join(pr, "r2_tuned_tiles.txt"))
    shutil.copy(os.path.join(go, tag + "_pytest.log"), os.path.join(pr, "r2_pytest_gpu.log"))
    grouped = subprocess.run([sys.executable, ext, os.path.join(go, tag + "_launches.csv"), "--group"], capture_output=True, text=True, check=True).stdout
    open(os.path.join(pr, "r2_launches_grouped.txt"), "w").write(grouped)
    for line in open(os.path.join(go, tag + "_bench.json")):
        if line.startswith("{"):
            json.dump(json.loads(line), open(os.path.join(pr, "r2_bench_n1.json"), "w"), indent=1)
    head = ("# round 2 (final kernels) -- ncu --set full --clock-control none captures, tile choices pinned to those of the un-profiled bench run\n"
            "# (PTK_TC_TUNE_FILE = profiles/r2_tuned_tiles.txt); first 1..8 launches of each family in one training step at 256x256 batch 8.\n"
            "# Cold-cache, serialised replays: read shares and ratios, not absolutes.  Families not re-captured here (gn_apply, gn_bwd_apply,\n"
            "# adam, splitk_reduce -- kernels unchanged since) are in r2a_ncu_summary.txt.\n")
    with open(os.path.join(pr, "r2_ncu_summary.txt"), "w") as f:
        f.write(head)
        for k in ("warp_forward_tiles", "warp_backward_tiles", "mask_pyramid", "nnloss_forward", "nnloss_backward", "gn_bwd_reduce",
                  "conv_tc_persist", "conv_tc_kernel", "wgrad_tc"):
            p = os.path.join(go, "%s_full_%s.csv" % (tag, k))
            if os.path.isfile(p):
                f.write("\n## %s\n" % k)
                f.write(subprocess.run([sys.executable, ext, p], capture_output=True, text=True, check=True).stdout)
    fam, tot, n = {}, 0.0, 0
    for line in grouped.splitlines():
        m = re.search(r"^(.*?)\s+n=(\d+)\s+us=([\d.]+)", line)
        if not m:
            continue
        name = next((v for k, v in FAMILIES if k in m.group(1)), "other")
        a = fam.setdefault(name, [0, 0.0])
        a[0] += int(m.group(2))
        a[1] += float(m.group(3))
        tot += float(m.group(3))
        n += int(m.group(2))
    print("launches %d, serialised kernel time %.1f us" % (n, tot))
    for k, (c, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print("%-28s n=%-4d us=%-9.1f share=%.3f" % (k, c, t, t / tot))


if __name__ == "__main__":
    main()
