"""One table of the headline ncu --set full metrics per kernel of a report (for profiles/):
python scripts/ncu_kernel_table.py gpurun_out/prof_kernels.ncu-rep > profiles/r01_kernels_ncu_full.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
COLS = [("gpu__time_duration.sum", "dur_us", 1.0),
        ("dram__bytes_read.sum", "dram_rd_MB", 1e-6), ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", 1.0),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu(mufu)_%", 1.0),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%", 1.0),
        ("launch__registers_per_thread", "regs", 1.0)]
print("source: %s   (ncu --set full --clock-control none; cold caches, serialised launches)" % rep)
print("%-44s %-14s " % ("kernel", "grid") + " ".join("%11s" % c[1] for c in COLS))
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    name = name.replace("capdec::", "").replace("<unnamed>::", "").replace("void ", "")
    name = name.split("(")[0][:44]
    vals = []
    for key, _, scale in COLS:
        if key not in idx:
            vals.append("-")
            continue
        v = float(r[idx[key]].replace(",", "")) * UNIT.get(units[idx[key]], 1)
        if key == "gpu__time_duration.sum":
            v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx[key]], 1.0)
        vals.append("%.1f" % (v * scale))
    print("%-44s %-14s " % (name, r[idx["Grid Size"]].replace(" ", "")) + " ".join("%11s" % v for v in vals))
