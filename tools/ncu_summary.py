#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv): per launch duration, IPC, occupancy, stall mix, memory."""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "dur"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.per_cycle_active", "warps/SM"), ("smsp__inst_executed.sum", "inst"),
        ("sm__inst_executed.avg.per_cycle_active", "ipc_act"), ("sm__inst_executed.avg.per_cycle_elapsed", "ipc_el"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__icc_request_hit_rate.pct", "icc_hit%")]
STALL = "smsp__average_warps_issue_stalled_"
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print("==", d["ID"], d["Kernel Name"][:60], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        print("  ", "  ".join(f"{n}={d[k]}{u[k] if u[k] not in ('', 'inst', 'warp') else ''}" for k, n in KEYS if k in d))
        st = {k[len(STALL):].replace("_per_issue_active.ratio", ""): float(v.replace(",", "")) for k, v in d.items() if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v}
        tot = sum(st.values())
        print("   stalls/issue:", "  ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]), f" total={tot:.2f}")
if __name__ == "__main__":
    main(sys.argv[1])
