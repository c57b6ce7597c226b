"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals over the LAST step.
usage: python tools/launch_summary.py launches.csv [n_steps_in_capture]"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    m = re.match(r"(?:void )?([\w:]+)(<.*)?\(", name)
    base = m.group(1) if m else name[:60]
    tm = re.search(r"<([^()]*)>\(", name)
    targs = tm.group(1) if tm and len(tm.group(1)) < 40 else ""
    if base.startswith("at::") or "elementwise" in name:
        f = re.search(r"at::native::(\w+)", name)
        base = "torch:" + (f.group(1) if f else base.split("::")[-1])
        fill = re.search(r"(FillFunctor|CUDAFunctor_add|MulFunctor|direct_copy)", name)
        if fill:
            base += ":" + fill.group(1)
        targs = ""
    return base + (f"<{targs}>" if targs else "")


def main():
    path = sys.argv[1]
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = len(rows) // nsteps
    last = rows[-per:]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in last:
        k = short(r["Kernel Name"])
        tot[k][0] += 1
        tot[k][1] += float(r["Metric Value"]) / 1e6
    total = sum(v[1] for v in tot.values())
    print(f"{len(rows)} launches in capture, {per} in the last step, {total:.2f} ms summed kernel time (serialised, cold cache)")
    print(f"{'kernel':70s} {'n':>5s} {'ms':>9s} {'share':>7s}")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:5d} {ms:9.3f} {100 * ms / total:6.1f}%")


if __name__ == "__main__":
    main()
