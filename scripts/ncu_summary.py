"""Reads the `--page raw --csv` exports of the round's ncu captures (gpurun_out/r2_prof_*_raw.csv) and writes
profiles/r2_ncu_summary.json + the stamped profiles/traffic.json that bench.py reads (CPU only)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "xbar2l1tex_read",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "pipe_tensor_active_pct",
    "sm__cycles_elapsed.max": "sm_cycles",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__cluster_size": "cluster",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}
UNITS = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0,
         "nsecond": 1e-9, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
CAPS = [("a5", "rmat16_a5", "bf16", "spmm_vbr_sm100 (pairs, fixed slots), BASELINE config #3"),
        ("er14", "er14_fixed", "bf16", "spmm_vbr_sm100, BASELINE config #2"),
        ("a4_tc", "rmat16_a4", "tf32", "spmm_vbr_sm100: the tensor-core part of the hybrid handle (block-rows taller than 7 rows)"),
        ("a4_gather", "rmat16_a4", "tf32", "spmm_csr_sm100: the gather part of the hybrid handle (block-rows of at most 7 rows)"),
        ("a4_tc_only", "rmat16_a4", "tf32", "spmm_vbr_sm100 with gather_max_height = -1: every block-row on the tensor cores")]


def read(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
    for i, h in enumerate(hdr):
        if h in WANT:
            try:
                out[WANT[h]] = float(vals[i].replace(",", "")) * UNITS.get(units[i], 1.0)
            except ValueError:
                pass
    return out


def main():
    summary, traffic = {}, {}
    sha = bench.source_sha16()
    for name, wl, prec, what in CAPS:
        p = os.path.join(ROOT, "gpurun_out", f"r2_prof_{name}_raw.csv")
        if not os.path.exists(p):
            continue
        r = read(p)
        r["what"] = what
        r["workload"] = f"{wl}:{prec}"
        if "dram_read" in r:
            r["dram_bytes"] = r["dram_read"] + r.get("dram_write", 0.0)
        summary[name] = r
        print(name, json.dumps(r))
    # the dominant kernel of each workload, stamped with the hash of the kernel sources it was taken from
    for name, key in (("a5", "rmat16_a5:bf16"), ("er14", "er14_fixed:bf16"), ("a4_tc", "rmat16_a4:tf32")):
        if name in summary:
            traffic[key] = {"dram_bytes": summary[name].get("dram_bytes"),
                            "pipe_tensor_active_pct": summary[name].get("pipe_tensor_active_pct"),
                            "xbar2l1tex_read_bytes": summary[name].get("xbar2l1tex_read"), "source_sha16": sha,
                            "source": f"profiles/r2_ncu_summary.json [{name}]: ncu --set full --clock-control none, one launch of "
                                      f"{summary[name].get('kernel')}"}
    json.dump(summary, open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json"), "w"), indent=1)
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print("source_sha16", sha)


if __name__ == "__main__":
    main()
