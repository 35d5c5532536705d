"""Key counters of `ncu --set full` reports as a markdown table.  usage: python scripts/ncu_summary.py a.ncu-rep ..."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (TPC)"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->tensor wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem lsu wavefronts %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__grid_size", "grid"),
    ("sm__cycles_elapsed.avg.per_second", "sm clk"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
]


def main():
    print("| report | kernel | " + " | ".join(k[1] for k in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"| {path} | (unreadable) |")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = d.get("Kernel Name", "?").split("(")[0].replace("void ", "").replace("sunet::", "")
            cells = []
            for k, _ in KEYS:
                v = d.get(k)
                cells.append("-" if v in (None, "") else f"{v} {u.get(k, '')}".strip())
            stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                             if "issue_stalled" in k and k.endswith("_per_warp_active.pct") and v not in ("", None)),
                            reverse=True)[:3]
            top = "; ".join(f"{k.split('issue_stalled_')[1].split('_per_warp')[0]} {x:.0f}%" for x, k in stalls)
            print(f"| {path.split('/')[-1]} | {name} | " + " | ".join(cells) + " |" + (f" stalls: {top}" if top else ""))


if __name__ == "__main__":
    main()
