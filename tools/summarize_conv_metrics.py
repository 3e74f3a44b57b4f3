"""Summarise `ncu --metrics ... --csv` launch lists of the convolution kernels (tools/train_step_ncu_target.py) into a table.
usage: python tools/summarize_conv_metrics.py "title" file.csv ["title" file.csv ...]"""
import collections, csv, sys

SHORT = (("conv_igemm_persistent_kernel", "persistent"), ("conv_wgrad_halo_kernel", "wgrad_halo"), ("conv_wgrad_kernel", "wgrad_splitk"),
         ("conv_halo_kernel", "halo"), ("conv_first_wgrad_kernel", "first_wgrad"), ("conv_first_kernel", "first"),
         ("pack_conv_weights_kernel", "pack_weights"))
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    out = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(h, r))
        e = out.setdefault(d["ID"], {"kernel": d["Kernel Name"], "grid": d.get("Grid Size", "")})
        e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
        e["unit_" + d["Metric Name"]] = d["Metric Unit"]
    return out


def main():
    args = sys.argv[1:]
    for title, path in zip(args[0::2], args[1::2]):
        L = load(path)
        print(f"== {title}: {len(L)} launches matching conv_ (ncu --clock-control none; per-launch times are serialised and cold-cache)")
        tot = tw = tens = 0.0
        for d in L.values():
            t, u = d["gpu__time_duration.sum"], d["unit_gpu__time_duration.sum"]
            t_us = t / 1000 if u.startswith("n") else (t if u.startswith("u") else t * 1000)
            tp = d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
            tc = d["l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
            gb = (d["dram__bytes_read.sum"] * BYTES[d["unit_dram__bytes_read.sum"]] +
                  d["dram__bytes_write.sum"] * BYTES[d["unit_dram__bytes_write.sum"]]) / 1e9
            kn = d["kernel"]
            for a, b in SHORT:
                kn = kn.replace(a, b)
            kn = kn.replace("void ", "").replace("unnamed>::", "").split("(")[0]
            if t_us < 60 and tp == 0:
                continue   # weight re-packing launches (5-50 us)
            print(f"  {kn[:22]:22s} grid {d['grid']:>13s} {t_us:8.1f} us  tensor pipe {tp:5.1f} %  operand reads (tc smem wavefronts) {tc:5.1f} %  dram {gb:5.2f} GB")
            if tp > 0:
                tens += t_us
                tw += t_us * tp
            tot += t_us
        print(f"  listed {tot / 1000:.2f} ms, of which tensor-core kernels {tens / 1000:.2f} ms at a time-weighted tensor pipe of {tw / max(tens, 1e-9):.1f} %")


if __name__ == "__main__":
    main()
