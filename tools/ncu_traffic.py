"""Turns `ncu -i <rep> --page raw --csv` of ONE res8 train step into profiles/r02_traffic.json (what bench.py reports as `roofline.traffic`):
per kernel label the average dram__bytes_read.sum + dram__bytes_write.sum per launch, the step total, and a markdown summary with duration,
DRAM throughput %, tensor-pipe activity and registers.   python tools/ncu_traffic.py raw.csv profiles/r02"""
import csv, json, sys, collections

LABELS = [("conv3x3_stream_tc_kernel<0", "conv3x3_fwd_tc"), ("conv3x3_stream_tc_kernel<1", "conv3x3_fwd_tc"), ("conv3x3_stream_tc_kernel<2", "conv3x3_dgrad_tc"),
          ("conv3x3_stream_tc_kernel<3", "conv3x3_dgrad_tc"), ("conv3x3_wgrad_tc_kernel", "conv3x3_wgrad_tc"), ("frontend_kernel", "frontend"),
          ("conv0_pool_kernel", "conv0_pool"), ("conv0_tc_kernel", "conv0_tc"), ("conv0_halo_kernel", "conv0_halo"), ("lstm_fwd_kernel", "lstm_fwd"), ("lstm_fwd_pipe_kernel", "lstm_fwd"), ("lstm_bwd_pipe_kernel", "lstm_bwd"),
          ("lstm_bwd_kernel", "lstm_bwd"), ("lstm_atb_kernel", "lstm_atb"), ("ctc_kernel", "ctc"), ("conv0_bwd_kernel", "conv0_bwd"), ("bn_bwd_head_op_kernel", "bn_bwd_head_op")]
WANT = {"dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr", "gpu__time_duration.sum": "ns", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct", "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
        "launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(path, out_prefix):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    kcol = col["Kernel Name"]
    per = collections.defaultdict(list)
    for r in rows[hdr + 2:]:
        if len(r) <= kcol:
            continue
        label = next((lab for pat, lab in LABELS if pat in r[kcol]), r[kcol][:40])
        # template kernels appear as conv3x3_stream_tc_kernel<(int)3, (bool)0>: normalise
        kn = r[kcol].replace("(int)", "").replace("(bool)", "").replace(" ", "")
        label = next((lab for pat, lab in LABELS if pat in kn), label)
        rec = {}
        for metric, key in WANT.items():
            if metric in col and r[col[metric]] not in ("", "n/a"):
                u = units[col[metric]]
                rec[key] = to_bytes(r[col[metric]], u) if key in ("rd", "wr") else float(r[col[metric]].replace(",", ""))
                if key == "ns":
                    rec[key] *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1)
        per[label].append(rec)
    traffic, lines, total = {}, ["| kernel | launches | avg us | DRAM bytes / launch | DRAM % of peak | tensor pipe % active | SM % | regs |", "|---|---|---|---|---|---|---|---|"], 0.0
    for label, recs in per.items():
        avg = lambda k: sum(r.get(k, 0.0) for r in recs) / len(recs)
        b = avg("rd") + avg("wr")
        traffic[label] = b
        total += sum(r.get("rd", 0.0) + r.get("wr", 0.0) for r in recs)
        lines.append(f"| {label} | {len(recs)} | {avg('ns') / 1e3:.1f} | {b / 1e6:.1f} MB | {avg('dram_pct'):.1f} | {avg('tensor_pct'):.1f} | {avg('sm_pct'):.1f} | {avg('regs'):.0f} |")
    traffic["__step_total__"] = total
    json.dump(traffic, open(out_prefix + "_traffic.json", "w"), indent=1)
    open(out_prefix + "_ncu_summary.md", "a").write("\n".join(lines) + f"\n\nDRAM bytes over the captured launches of one step: {total / 1e9:.2f} GB\n")
    print("\n".join(lines)); print("total GB", total / 1e9)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
