"""Key metrics of one profiled launch from an `ncu --set full` report, as the small JSON record committed under profiles/.
Usage: python tools/ncu_summary.py REPORT.ncu-rep "<command that was profiled>" [batch_per_launch] [precision] > profiles/NAME.json
(`ncu -i REPORT --page raw --csv` is run here; the record with "dram_bytes_per_launch" also feeds bench.py's roofline.traffic)"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3,
        "Ghz": 1.0, "Mhz": 1e-3, "hz": 1e-9}

WANT = {
    "duration_ms_under_ncu": "gpu__time_duration.sum",
    "sm_clock_ghz_under_ncu": "sm__cycles_elapsed.avg.per_second",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tc_smem_wavefronts_pct_of_peak": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "issue_slots_busy_pct": "sm__inst_executed.avg.pct_of_peak_sustained_elapsed",
    "lsu_wavefronts_pct_of_peak": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "l1_read_bytes_from_l2": "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1_write_bytes_to_l2": "l1tex__m_l1tex2xbar_write_bytes.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "block_size": "launch__block_size",
    "grid_size": "launch__grid_size",
    "dyn_smem_per_block_bytes": "launch__shared_mem_per_block_dynamic",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "warp_instructions": "smsp__inst_executed.sum",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(head)}
    out = {"kernel": vals[col["Kernel Name"]].replace("void ", "").split("(")[0],
           "source": "%s (ncu --set full --clock-control none; %s)" % (rep.split("/")[-1], sys.argv[2] if len(sys.argv) > 2 else ""),
           "batch_per_launch": int(sys.argv[3]) if len(sys.argv) > 3 else 8192,
           "precision": sys.argv[4] if len(sys.argv) > 4 else "bf16"}
    for k, name in WANT.items():
        if name not in col or vals[col[name]] in ("", "n/a"):
            continue
        v = float(vals[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)
        out[k] = v
    if "dram_bytes_read" in out and "dram_bytes_write" in out:
        out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
        out["dram_gb_per_s_under_ncu"] = out["dram_bytes_per_launch"] / (out["duration_ms_under_ncu"] * 1e-3) / 1e9
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
