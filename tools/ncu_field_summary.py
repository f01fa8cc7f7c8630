"""ncu --set full report of field_forward launches + the rows of those launches -> profiles/r02_field_kernel_ncu.json
(read by bench.py for roofline.traffic / roofline.ncu).  Usage (here, after the report came back in gpurun_out/):
    python tools/ncu_field_summary.py gpurun_out/r2_final_field.ncu-rep gpurun_out/field_rows.json <first launch index>"""
import csv
import io
import json
import subprocess
import sys

rep, rows_json, first = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
t = list(csv.reader(io.StringIO(raw)))
hdr, data = t[0], t[2:]
col = lambda name: [float(r[hdr.index(name)].replace(",", "")) for r in data]
unit = lambda name: t[1][hdr.index(name)]
rows = [r["rows"] for r in json.load(open(rows_json))][first:first + len(data)]
to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = [v * to_bytes[unit("dram__bytes_read.sum")] for v in col("dram__bytes_read.sum")]
wr = [v * to_bytes[unit("dram__bytes_write.sum")] for v in col("dram__bytes_write.sum")]
us = [v * {"us": 1, "ms": 1e3, "ns": 1e-3}[unit("gpu__time_duration.sum")] for v in col("gpu__time_duration.sum")]
n = sum(rows)
out = {
    "source": f"{rep} (ncu --set full --clock-control none, launches {first}..{first + len(data) - 1} of tools/field_profile_target.py)",
    "launch_rows": rows, "launch_us": us,
    "dram_bytes_per_row": (sum(rd) + sum(wr)) / n,
    "dram_read_bytes_per_row": sum(rd) / n, "dram_write_bytes_per_row": sum(wr) / n,
    "grows_per_s_under_ncu": n / (sum(us) * 1e-6) / 1e9,
    "lts_throughput_pct": sum(col("lts__throughput.avg.pct_of_peak_sustained_elapsed")) / len(data),
    "lts_sector_hit_rate_pct": sum(col("lts__t_sector_hit_rate.pct")) / len(data),
    "l1tex_sector_hit_rate_pct": sum(col("l1tex__t_sector_hit_rate.pct")) / len(data),
    "l1tex_data_pipe_lsu_wavefronts_pct": sum(col("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")) / len(data),
    "smsp_issue_active_pct": sum(col("smsp__issue_active.avg.pct_of_peak_sustained_active")) / len(data),
    "sm_warps_active_pct": sum(col("sm__warps_active.avg.pct_of_peak_sustained_active")) / len(data),
    "tensor_pipe_active_pct": sum(col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")) / len(data),
    "warp_instructions_per_row": sum(col("smsp__inst_executed.sum")) / n,
    "registers_per_thread": col("launch__registers_per_thread")[0],
}
json.dump(out, open("profiles/r02_field_kernel_ncu.json", "w"), indent=1)
print(json.dumps(out, indent=1))
