"""Print the counters that matter from an .ncu-rep (ncu -i ... --page raw --csv): duration, issue slots, tensor pipe, stall split,
registers, occupancy, DRAM traffic.   python tools/ncu_summary.py gpurun_out/x.ncu-rep [row]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + row]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][0]} {d[k][1]}")
stalls = sorted(((float(v[0]), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")), reverse=True)
print("stall reasons (warps stalled per issued instruction):")
for v, k in stalls[:12]:
    print(f"   {k.split('stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
