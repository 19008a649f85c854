"""profiles/<tag>_ncu_full_summary.json from the `ncu -i ... --page raw --csv` exports of one GPU call (gpurun_out/<tag>_ncu_full_*.raw.csv)."""
import csv
import glob
import json
import os
import sys

tag = sys.argv[1]
want = {
    "ms": ("gpu__time_duration.sum", 1e-6),   # reported in ns (or us/ms: the unit row is applied below)
    "dram_read_bytes": ("dram__bytes_read.sum", 1.0),
    "dram_write_bytes": ("dram__bytes_write.sum", 1.0),
    "dmma_pipe_pct": ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", 1.0),
    "tensor_fp64_ops_pct_of_peak_elapsed": ("sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", 1.0),
    "issue_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "shared_bank_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0),
    "registers_per_thread": ("launch__registers_per_thread", 1.0),
}
unit_scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {}
for f in sorted(glob.glob(f"gpurun_out/{tag}_ncu_full_*.raw.csv")):
    rows = list(csv.reader(open(f)))
    hdr, units, val = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    rec = {"kernel": val[col["Kernel Name"]][:110]}
    for key, (name, _) in want.items():
        if name in col and val[col[name]] not in ("", "n/a"):
            x = float(val[col[name]].replace(",", ""))
            u = units[col[name]]
            if key == "ms" or key.endswith("_bytes"):
                x *= unit_scale.get(u, 1.0)
            rec[key] = x
    out[os.path.basename(f)[len(tag) + len("_ncu_full_"):-len(".raw.csv")]] = rec
json.dump(out, open(f"profiles/{tag}_ncu_full_summary.json", "w"), indent=1)
print(json.dumps(out, indent=1))
