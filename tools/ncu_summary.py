"""Summarise an `ncu --set full` report (.ncu-rep) per kernel: duration, DRAM traffic, L2/L1 throughput, occupancy,
registers, issue-slot utilisation, top warp-stall reasons. Usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe insts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("=" * 100)
        print("kernel:", d["Kernel Name"][:120])
        for key, name in WANT:
            if key in d:
                print(f"  {name:28s} {d[key]:>16s} {u[key]}")
        stalls = [(float(d[k].replace(",", "")), k) for k in hdr
                  if k.startswith("smsp__average_warp_latency_issue_stalled") and k.endswith("_per_warp_active.pct") is False
                  and "ratio" in k and d[k] not in ("", "n/a")]
        stalls2 = [(float(d[k].replace(",", "")), k) for k in hdr
                   if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and d[k] not in ("", "n/a")]
        for v, k in sorted(stalls2, reverse=True)[:6]:
            print(f"  stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:10.2f} warps/issue")


if __name__ == "__main__":
    main(sys.argv[1])
