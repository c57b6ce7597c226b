"""Summarise .ncu-rep captures (ncu --set full) as a markdown table: duration, DRAM bytes, achieved GB/s and fraction of the measured HBM
peak, occupancy, tensor-pipe activity, registers, top warp-stall reasons.   usage: python tools/ncu_summary.py out.md rep1.ncu-rep [...]
(test / measurement infrastructure; reads reports brought back from the GPU box, runs without a GPU)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6550.1


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def stalls(rep, kernel_index):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    if kernel_index >= len(blocks) or not blocks[kernel_index]:
        return ""
    hdr, data = blocks[kernel_index][0], blocks[kernel_index][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    agg = {}
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            agg[h[6:]] = sum(int(r[ix[h]] or 0) for r in data)
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
    return ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top)


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    hbm = peaks()
    lines = ["| capture | kernel | grid | duration us | DRAM read + write MB | DRAM GB/s (frac of measured " + f"{hbm:.0f}" +
             ") | warps active % | tensor pipe % | regs | top warp stalls |", "|---|---|---|---|---|---|---|---|---|---|"]
    for rep in reps:
        for i, d in enumerate(raw(rep)):
            g = lambda k, default="": d.get(k, default)
            us = float(g("gpu__time_duration.sum", "0") or 0)
            rd, wr = float(g("dram__bytes_read.sum", "0") or 0), float(g("dram__bytes_write.sum", "0") or 0)
            unit = 1.0          # ncu prints MB for these sizes; GB for the arena-sized kernels
            if rd + wr < 20:    # GB
                unit = 1000.0
            mb = (rd + wr) * unit
            gbs = mb / 1e3 / (us / 1e6) if us else 0.0
            lines.append(f"| {os.path.basename(rep)} | {g('Kernel Name')[:52]} | {g('Grid Size')} | {us:.1f} | {rd * unit:.1f} + {wr * unit:.1f} | "
                         f"{gbs:.0f} ({gbs / hbm:.2f}) | {float(g('sm__warps_active.avg.pct_of_peak_sustained_active', '0') or 0):.0f} | "
                         f"{float(g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', '0') or 0):.0f} | "
                         f"{g('launch__registers_per_thread')} | {stalls(rep, i)} |")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
