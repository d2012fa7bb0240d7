"""Extract per-launch DRAM traffic of a kernel from an `ncu --set full` report and write the small JSON that
bench.py's roofline object reads (profiles/r01_head_gemm_ncu.json).
usage: python scripts/ncu_traffic.py gpurun_out/prof_head.ncu-rep 'gemm_tc_kernel<256, 0, 1' profiles/r01_head_gemm_ncu.json"""
import csv
import io
import json
import subprocess
import sys

rep, pattern, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
sel = [r for r in rows[2:] if pattern in r[idx["Kernel Name"]]]


def val(r, name):
    return float(r[idx[name]].replace(",", "")) * UNIT.get(units[idx[name]], 1)


res = {"kernel": pattern, "launches": len(sel),
       "dram_bytes_per_launch": sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in sel) / max(1, len(sel)),
       "dram_read_bytes_per_launch": sum(val(r, "dram__bytes_read.sum") for r in sel) / max(1, len(sel)),
       "dram_write_bytes_per_launch": sum(val(r, "dram__bytes_write.sum") for r in sel) / max(1, len(sel)),
       "duration_us_under_ncu": sum(float(r[idx["gpu__time_duration.sum"]].replace(",", "")) for r in sel) / max(1, len(sel)),
       "source": rep}
for name in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
             "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
    if name in idx and sel:
        res[name] = sum(float(r[idx[name]].replace(",", "")) for r in sel) / len(sel)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
