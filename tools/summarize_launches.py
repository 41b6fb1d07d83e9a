"""ncu launch list (gpu__time_duration.sum per launch, --csv) -> per-kernel summary CSV."""
import collections, csv, sys
src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    name = r[ik].split("(")[0][:90]
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
with open(dst, "w") as f:
    f.write(f"# summary of {src}: {sum(cnt.values())} launches, {s/1e3:.1f} us total (cold-cache, serialised: compare shares)\n")
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us"])
    for k, v in tot.most_common():
        w.writerow([k, cnt[k], f"{v/1e3:.2f}", f"{100*v/s:.2f}", f"{v/cnt[k]/1e3:.2f}"])
