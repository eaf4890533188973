#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the tables kept under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv        # per-kernel totals + shares (markdown)
  python profiles/summarize.py kernels  gpurun_out/prof.ncu-rep [...]  # key `--set full` metrics per captured kernel

The launch list is the `ncu --metrics gpu__time_duration.sum --clock-control none --csv` pass of
/opt/skills/guides/B200_PROFILING.md; its per-launch times are cold-cache and serialised, so only the SHARES are
compared with the CUDA-event stage times of the bench line.
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    tot = OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        ms = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6) * v
        name = r[ik].split("(")[0]
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in tot.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        if ms / total < 0.001:
            continue
        print("| `%s` | %d | %.3f | %.1f%% |" % (name[:80], n, ms, 100 * ms / total))
    own = sum(ms for k, (n, ms) in tot.items() if "dgs::" in k or k.startswith("dgs"))
    cub = sum(ms for k, (n, ms) in tot.items() if "cub::" in k)
    print("\nOwn kernels (`dgs::*`): %.1f %% of GPU time; CUB scan/sort: %.1f %%; torch elementwise/reduce: %.1f %%."
          % (100 * own / total, 100 * cub / total, 100 * (total - own - cub) / total))


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_allocated", "smem / block"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "global RED sectors"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe (LSU) wavefronts, % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles active %"),
]
STALLS = ["barrier", "wait", "short_scoreboard", "long_scoreboard", "not_selected", "selected", "branch_resolving",
          "math_pipe_throttle", "lg_throttle", "mio_throttle", "dispatch_stall", "no_instruction"]


def kernels(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print("## %s\n" % path)
        names = [r[hdr.index("Kernel Name")].split("(")[0] for r in rows[2:]]
        print("| metric | " + " | ".join("`%s`" % n for n in names) + " |\n|---|" + "---|" * len(names))
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("| %s (%s) | " % (label, units[i]) + " | ".join(r[i] for r in rows[2:]) + " |")
        tot = [0.0] * len(names)
        vals = {}
        for s in STALLS:
            key = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if key in hdr:
                i = hdr.index(key)
                vals[s] = [float(r[i]) for r in rows[2:]]
                tot = [a + b for a, b in zip(tot, vals[s])]
        for s, v in vals.items():
            print("| stall %s (%% of warp cycles) | " % s + " | ".join("%.1f" % (100 * a / t) for a, t in zip(v, tot)) + " |")
        print()


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "kernels"):
        sys.exit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernels(sys.argv[2:])
