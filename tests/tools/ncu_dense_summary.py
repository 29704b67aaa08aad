"""Summarise an `ncu --set full` capture of the dense conv kernel: per captured launch the grid, duration, DRAM bytes,
tensor-pipe activity.  Usage: python tests/tools/ncu_dense_summary.py gpurun_out/prof_r2_dense.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "smsp__cycles_active.avg"]
print("| # | kernel | grid | us | DRAM rd | DRAM wr | L2->SM | tensor pipe active % |")
print("|---|---|---|---:|---:|---:|---:|---:|")
for n, r in enumerate(rows[2:]):
    def g(k):
        return r[col[k]] if k in col else "?"
    def gu(k):
        return (r[col[k]] + " " + units[col[k]]) if k in col else "?"
    print("| %d | `%s` | %s | %s | %s | %s | %s | %s |" % (
        n, g("Kernel Name").split("(")[0][-40:], g("Grid Size"), g("gpu__time_duration.sum"), gu("dram__bytes_read.sum"),
        gu("dram__bytes_write.sum"), gu("l1tex__m_xbar2l1tex_read_bytes.sum"),
        g("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")))
