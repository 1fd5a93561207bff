"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into profiles/<name>.json/.md."""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        # per-slice balance of the L2 (avg / max / min over the slices)
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.min.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.avg", "lts__t_sectors.max", "lts__t_sectors.min",
        "lts__throughput.max.pct_of_peak_sustained_elapsed",
        "lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__xbar2lts_cycles_active.max.pct_of_peak_sustained_elapsed",
        "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_requests.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        # issue and pipes (ALU-bound kernels)
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val) * mult.get(unit, 1)


def main(rep, out_prefix, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": vals[i], "unit": units[i]}
        stalls = []
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                stalls.append((float(vals[i]), h.split("stalled_")[1].split("_per_issue")[0]))
        d["top_stalls_per_issue"] = [{"reason": r, "ratio": round(v, 2)} for v, r in sorted(stalls, reverse=True)[:6]]
        rd = d.get("dram__bytes_read.sum")
        wr = d.get("dram__bytes_write.sum")
        if rd and wr:
            d["dram_traffic_bytes"] = to_bytes(rd["value"], rd["unit"]) + to_bytes(wr["value"], wr["unit"])
        out.append(d)
    json.dump({"report": rep, "note": note, "launches": out}, open(out_prefix + ".json", "w"), indent=1)
    with open(out_prefix + ".md", "w") as f:
        f.write("# ncu summary: %s\n\n%s\n\n" % (rep, note))
        for d in out:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d["kernel"])
            for k in KEYS:
                if k in d:
                    f.write("| %s | %s | %s |\n" % (k, d[k]["value"], d[k]["unit"]))
            if "dram_traffic_bytes" in d:
                f.write("| dram traffic (read+write) | %.4g | byte |\n" % d["dram_traffic_bytes"])
            f.write("\nTop stall reasons (warps stalled per issue-active cycle): " +
                    ", ".join("%s %.2f" % (s["reason"], s["ratio"]) for s in d["top_stalls_per_issue"]) + "\n\n")
    print("wrote", out_prefix + ".json/.md")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
