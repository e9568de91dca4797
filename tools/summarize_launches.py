"""ncu launch-list CSV (--metrics gpu__time_duration.sum --csv) -> per-kernel totals as JSON (profiles/)."""
import collections
import csv
import json
import sys


def summarize(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    cols = rows[hdr]
    ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 2:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki]
        short = name.split("(")[0]
        if "<" in name and ("gemm" in name or "attn" in name):
            short = name.split("(")[0] + "<" + name.split("<", 1)[1].split(">(")[0][:60] + ">"
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = {"total_us": tot / 1e3, "launches": sum(a[0] for a in agg.values()),
           "kernels": [{"kernel": k, "launches": a[0], "total_us": round(a[1] / 1e3, 1), "avg_us": round(a[1] / a[0] / 1e3, 2),
                        "share": round(a[1] / tot, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    return out


if __name__ == "__main__":
    s = summarize(sys.argv[1])
    if len(sys.argv) > 2:
        json.dump(s, open(sys.argv[2], "w"), indent=1)
    for k in s["kernels"][:30]:
        print("%5d  %10.1f us  %5.1f%%  avg %8.2f us  %s" % (k["launches"], k["total_us"], 100 * k["share"], k["avg_us"], k["kernel"][:110]))
    print("total", round(s["total_us"], 1), "us over", s["launches"], "launches")
