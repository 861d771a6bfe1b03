"""Per-kernel shares of ONE steady-state step from an ncu launch list.

  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline
  python tools/launch_shares.py launches.csv one_step.csv shares.txt

A step starts at the weight bank's statistics kernel (alignq::wq_stats_kernel); the LAST complete step of the run is
taken.  ncu serialises the kernels and profiles them cold, so the side-stream kernels that overlap the main chain in the
real step are counted in full: compare SHARES, not the sum, with the CUDA-event time of bench.py."""
import collections
import csv
import sys

src, one_step, out = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
OWN = ("alignq", "ctc::", "head::", "stem::", "lmmd::", "tcsmall::")      # ncu prints the innermost namespaces only


def is_own(k):
    return any(t in k for t in OWN)


with open(src, newline="") as f:
    first = f.readline()
    f.seek(0)
    if first.startswith("#"):                 # already a one-step list written by this tool: recompute the summary
        rd0 = csv.reader(f)
        next(rd0)
        pre = [(r[1], float(r[2])) for r in rd0]
        lines = None
    else:
        lines = [l for l in f if l.startswith('"')]
if lines is None:
    rows = [("alignq::wq_stats_kernel", 0.0)] if not pre or "wq_stats" not in pre[0][0] else []
    rows = pre + [("alignq::wq_stats_kernel(end)", 0.0)]
    lines = ['"Kernel Name","Metric Name","Metric Value","Metric Unit"']
rd = csv.reader(lines)
hdr = next(rd)
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) > vi and r[mi] == "gpu__time_duration.sum":
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
        rows.append((r[ki], us))
starts = [i for i, (k, _) in enumerate(rows) if "wq_stats_kernel" in k]
assert len(starts) >= 2, "need at least two steps in the launch list"
pairs = [(starts[i], starts[i + 1]) for i in range(len(starts) - 1)]
fused = [(a, b) for a, b in pairs if any("bnq_apply_kernel" in k for k, _ in rows[a:b])]    # not the --no-fuse comparison model
a, b = (fused or pairs)[-1]
step = [(k, us) for k, us in rows[a:b] if not ("FillFunctor<unsigned char>" in k or "FillFunctor<signed char>" in k)]   # the L2 flush
with open(one_step, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["#", "kernel", "gpu__time_duration.sum [us]"])
    for i, (k, us) in enumerate(step):
        w.writerow([i, k, f"{us:.3f}"])
tot = sum(us for _, us in step)
agg = collections.OrderedDict()
for k, us in step:
    k = k.split("(")[0]
    t = agg.setdefault(k, [0.0, 0])
    t[0] += us
    t[1] += 1
own = sum(t[0] for k, t in agg.items() if is_own(k))
nown = sum(t[1] for k, t in agg.items() if is_own(k))
with open(out, "w") as f:
    f.write(f"# kernels in the step: {len(step)}; serialised sum {tot:.1f} us\n")
    f.write(f"# alignq_b200 kernels: {nown} launches, {own:.1f} us = {100 * own / tot:.1f}% of the serialised sum\n")
    f.write("#   us      share  launches  kernel\n")
    for k, t in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{t[0]:9.1f}  {100 * t[0] / tot:5.1f}% x{t[1]:4d}  {k[:110]}\n")
print(open(out).read()[:2500])
