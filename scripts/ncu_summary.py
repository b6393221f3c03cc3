"""Key counters of every launch in an `ncu --set full` report (read through `ncu -i X --page raw --csv`).
    python scripts/ncu_summary.py report.ncu-rep > profiles/summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global st sectors"),
]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(f"== {r[idx['Kernel Name']][:110]}  (launch id {r[idx['ID']]})")
    for key, label in WANT:
        if key in idx:
            print(f"   {label:24s} {r[idx[key]]:>18s} {units[idx[key]]:12s} {key}")
    stalls = [(h, r[i]) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    stalls = sorted(((h, float(v)) for h, v in stalls if v not in ("", "n/a")), key=lambda x: -x[1])[:5]
    print("   top stalls (warps per issue):", ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for h, v in stalls))
