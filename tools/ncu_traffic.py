"""Turn `ncu -i REP --page raw --csv` of one kernel launch into the small JSON record bench.py reads for
`roofline.traffic` (profiles/*traffic*.json).
Usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_traffic.py raw.csv "<source note>" > profiles/rN_traffic.json"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3,
        "Ghz": 1.0, "Mhz": 1e-3, "hz": 1e-9, "%": 1.0, "register/thread": 1.0}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    head, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(head)}

    def get(name):
        return float(vals[col[name]].replace(",", "")) * UNIT[units[col[name]]]

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    out = {
        "kernel": vals[col["Kernel Name"]].replace("void ", "").split("(")[0],
        "source": sys.argv[2] if len(sys.argv) > 2 else "",
        "batch_per_launch": int(sys.argv[3]) if len(sys.argv) > 3 else 8192,
        "dram_bytes_read": rd,
        "dram_bytes_write": wr,
        "dram_bytes_per_launch": rd + wr,
        "duration_ms_under_ncu": get("gpu__time_duration.sum"),
        "sm_clock_ghz_under_ncu": get("sm__cycles_elapsed.avg.per_second"),
        "tensor_pipe_active_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "tc_smem_wavefronts_pct_of_peak": get("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "l2_hit_rate_pct": get("lts__t_sector_hit_rate.pct"),
        "l1_read_bytes_from_l2": get("l1tex__m_xbar2l1tex_read_bytes.sum"),
        "l1_write_bytes_to_l2": get("l1tex__m_l1tex2xbar_write_bytes.sum"),
        "registers_per_thread": int(get("launch__registers_per_thread")),
    }
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
