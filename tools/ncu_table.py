"""print a compact per-kernel table from an .ncu-rep (raw page): duration, DRAM bytes, tensor/fp32 pipe, occupancy"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    for i, h in enumerate(hdr):
        if h == name: return i
    return None
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fmacyc%"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf")]
idx = [(col(n), lab) for n, lab in want if col(n) is not None]
print(" | ".join(lab + ("(" + units[i] + ")" if units[i] else "") for i, lab in idx))
for r in data:
    out = []
    for i, lab in idx:
        v = r[i]
        if lab == "kernel": v = v.split("(")[0][-45:]
        out.append(v)
    print(" | ".join(out))
