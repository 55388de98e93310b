"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name: count, mean ms, share."""
import collections
import csv
import sys


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    d = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[0].isdigit():
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        d.setdefault(r[ki], []).append(v)
    return d


if __name__ == "__main__":
    d = summarise(sys.argv[1])
    tot = sum(sum(v) for v in d.values())
    for k, v in d.items():
        print(f"{k[:84]:84s} n={len(v):3d} avg={sum(v) / len(v):8.4f} ms  share={100 * sum(v) / tot:5.1f}%")
    print(f"total {tot:.3f} ms over {sum(len(v) for v in d.values())} launches")
