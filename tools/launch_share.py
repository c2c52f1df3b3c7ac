"""Per-kernel share of a decode step from an `ncu --metrics gpu__time_duration.sum --csv` launch list (CPU only).

    python tools/launch_share.py gpurun_out/r2_decode_step_launches.csv > profiles/r2_decode_step_share.txt
ncu durations are cold-cache and serialised: the SHARE per kernel is what must agree with the bench, not the absolute time."""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 2:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r"b200q::", "", re.sub(r"\(.*", "", r[ki]))
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui].startswith("n") else (v if r[ui].startswith("u") else v * 1e3)
    tot[name] += us
    cnt[name] += 1
T = sum(tot.values())
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {T:.0f} us (ncu: cold cache, serialised)")
for k, v in tot.most_common(24):
    print(f"{100 * v / T:5.1f}%  {v:10.1f} us  n={cnt[k]:5d}  avg {v / cnt[k]:8.2f} us  {k[:120]}")
