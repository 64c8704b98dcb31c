"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one record per kernel launch with the
metrics DESIGN.md quotes.  python tools/ncu_full_summary.py raw.csv "case description" > out.json"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "launch__grid_size": "grid",
    "launch__registers_per_thread": "regs",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed.sum": "inst_executed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "sm__cycles_elapsed.max": "sm_cycles",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "sm__inst_executed_pipe_fp64.sum": "fp64_inst",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__inst_executed.sum": "warp_inst",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main(path, case):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    # newer ncu versions prefix some metrics with their section ("FBSP.TriageCompute.dram__..."): index by suffix too
    alias = {}
    for h in hdr:
        for k in WANT:
            if h == k or h.endswith("." + k):
                alias.setdefault(k, h)
    out = []
    for r in rows[2:]:
        rec = {"case": case}
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = re.sub(r"\(.*", "", d.get("Kernel Name", ""))
        rec["kernel"] = re.sub(r"void |slk::", "", name)
        for k0, short in WANT.items():
            k = alias.get(k0, k0)
            if k in d and num(d[k]) is not None:
                v = num(d[k])
                if short == "duration_us":
                    v = v / 1e3 if u[k] in ("ns", "nsecond") else (v * 1e3 if u[k] in ("ms", "msecond") else v)
                if short.endswith("_MB"):
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u[k], 1e-6)
                    v = v * scale
                rec[short] = round(v, 3)
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
