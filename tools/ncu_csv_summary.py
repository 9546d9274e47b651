"""Summarise an `ncu --page raw --csv` export: one block per kernel launch with the metrics the roofline discussion uses."""
import csv
import json
import sys

KEYS = [("us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"),
        ("waves_per_sm", "launch__waves_per_multiprocessor"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warp_inst", "smsp__inst_executed.sum"),
        ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor_pipe_pct", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
        ("tensor_pipe_cycles_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("eligible_warps_per_cycle", "smsp__warps_eligible.avg.per_cycle_active")]


def main(path, out=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for name, k in KEYS:
            if k in hdr and r[hdr.index(k)] != "":
                v = float(r[hdr.index(k)].replace(",", ""))
                u = units[hdr.index(k)]
                if name.endswith("_MB"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if name == "us":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
                d[name] = round(v, 3)
        stalls = []
        for i, h in enumerate(hdr):                      # warp-state samples (pc sampling), issued and not issued together
            if "pcsamp_warps_issue_stalled_" in h and "not_issued" not in h:
                try:
                    stalls.append((float(r[i].replace(",", "")), h.split("issue_stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        d["warp_state_pct"] = {n: round(100 * v / tot, 1) for v, n in sorted(stalls, reverse=True)[:6]}
        res.append(d)
    if out:
        json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(json.dumps(d))


if __name__ == "__main__":
    main(*sys.argv[1:3])
