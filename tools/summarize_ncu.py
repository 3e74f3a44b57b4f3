#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into profiles/: key metrics per launch + stall reasons.

usage: python tools/summarize_ncu.py gpurun_out/r1_rcps_hist.ncu-rep profiles/r1_rcps_hist --images 10000 --side 320
"""
import argparse, csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "sm__cycles_elapsed.avg", "lts__t_bytes.sum"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("out_prefix")
    ap.add_argument("--images", type=int, default=None); ap.add_argument("--side", type=int, default=None)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                try:
                    d[k] = float(v)
                except ValueError:
                    d[k] = v
                d[k + ".unit"] = units[hdr.index(k)]
        stalls = {}
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        d["stall_pct"] = {k: round(100 * v / tot, 2) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        launches.append(d)
    def to_bytes(d, k):
        u = d.get(k + ".unit", "")
        m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return d.get(k, 0) * m
    last = launches[-1]
    summary = {"report": a.rep, "images": a.images, "side": a.side, "launches": launches,
               "dram_bytes_per_launch": to_bytes(last, "dram__bytes_read.sum") + to_bytes(last, "dram__bytes_write.sum")}
    json.dump(summary, open(a.out_prefix + "_ncu_summary.json", "w"), indent=1)
    with open(a.out_prefix + "_ncu_summary.txt", "w") as f:
        for d in launches:
            f.write(d["kernel"] + "\n")
            for k in KEYS:
                if k in d:
                    f.write(f"  {k:75s} {d[k]!s:>22s} {d.get(k + '.unit', '')}\n")
            f.write(f"  stall reasons (% of samples): {d['stall_pct']}\n\n")
    print(open(a.out_prefix + "_ncu_summary.txt").read())


if __name__ == "__main__":
    main()
