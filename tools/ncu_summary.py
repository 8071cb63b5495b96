"""Summarise ncu output into the small text files kept under profiles/.

  python tools/ncu_summary.py launches <launch_list.csv> <out_prefix> [--pairs N]
      ncu --metrics gpu__time_duration.sum --clock-control none --csv  ->  <out_prefix>.csv (compact launch list)
      and <out_prefix>.md (per-kernel launches / total / share of the step)
  python tools/ncu_summary.py raw <raw_page.csv> <out.md>
      `ncu -i x.ncu-rep --page raw --csv`  ->  the metrics the roofline discussion uses, one block per kernel
"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"mb2_\w+_detail::|\(anonymous namespace\)::", "", name)
    return name


def read_csv(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def launches(path, prefix, pairs):
    rows = [r for r in read_csv(path) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = defaultdict(lambda: [0, 0.0])
    with open(prefix + ".csv", "w") as f:
        f.write("id,kernel,grid,block,ns\n")
        for r in rows:
            k = short(r["Kernel Name"]); ns = float(r["Metric Value"].replace(",", ""))
            agg[k][0] += 1; agg[k][1] += ns
            f.write('%s,"%s","%s","%s",%d\n' % (r["ID"], k, r["Grid Size"], r["Block Size"], ns))
    ours = {k: v for k, v in agg.items() if k.startswith("k_") or "cub" in k.lower() or "DeviceRadixSort" in k}
    tot = sum(v[1] for v in ours.values())
    with open(prefix + ".md", "w") as f:
        f.write("# ncu launch list summary (%s)\n\n" % path.split("/")[-1])
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none --csv` over `bench.py` (%d pairs in the capture: warm-up + timed + profile pass).\n" % pairs)
        f.write("Per-launch times under ncu are cold-cache and serialised: use the SHARE column, not the absolute.\n\n")
        f.write("| kernel | launches | launches/pair | total ms | ms/pair | share |\n|---|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.3f | %.3f | %.1f%% |\n" % (k, v[0], v[0] / pairs, v[1] / 1e6, v[1] / 1e6 / pairs, 100 * v[1] / tot))
        f.write("| **all library kernels** | %d | %.1f | %.3f | %.3f | 100%% |\n" % (sum(v[0] for v in ours.values()), sum(v[0] for v in ours.values()) / pairs, tot / 1e6, tot / 1e6 / pairs))
        other = {k: v for k, v in agg.items() if k not in ours}
        if other:
            f.write("\nOther kernels in the process (torch fills / copies made by bench.py itself): %d launches, %.3f ms.\n" % (sum(v[0] for v in other.values()), sum(v[1] for v in other.values()) / 1e6))


KEYS = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occ. limit (regs)"),
    ("launch__occupancy_limit_shared_mem", "occ. limit (smem)"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier (warps/issue)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (warps/issue)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected (warps/issue)"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction (warps/issue)"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving (warps/issue)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
])


def raw(path, out):
    rows = read_csv(path)
    with open(out, "w") as f:
        f.write("# ncu --set full, selected metrics (%s)\n\n" % path.split("/")[-1])
        f.write("One capture per kernel (`ncu --set full --clock-control none --import-source on`), read with `ncu -i ... --page raw --csv`.\n\n")
        units = rows[0] if rows and rows[0].get("ID", "") == "" else None
        for r in rows:
            if r is units:
                continue
            f.write("## %s  grid %s block %s\n\n" % (short(r.get("Kernel Name", "?")), r.get("Grid Size", "?"), r.get("Block Size", "?")))
            for k, label in KEYS.items():
                if k in r and r[k] != "":
                    f.write("- %s (`%s`): %s %s\n" % (label, k, r[k], units.get(k, "") if units else ""))
            try:
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                rd = float(r["dram__bytes_read.sum"].replace(",", "")) * mult[units["dram__bytes_read.sum"]]
                wr = float(r["dram__bytes_write.sum"].replace(",", "")) * mult[units["dram__bytes_write.sum"]]
                dur = float(r["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units["gpu__time_duration.sum"]]
                f.write("- **traffic (dram read+write)**: %.1f MB  (%.0f GB/s over the launch)\n" % ((rd + wr) / 1e6, (rd + wr) / 1e9 / dur))
            except Exception:
                pass
            f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        pairs = int(sys.argv[sys.argv.index("--pairs") + 1]) if "--pairs" in sys.argv else 1
        launches(sys.argv[2], sys.argv[3], pairs)
    else:
        raw(sys.argv[2], sys.argv[3])
