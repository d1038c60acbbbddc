"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel. Usage: python tools/summarize_launches.py file.csv [top]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    out = []
    for d in data:
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        out.append((re.sub(r"\(.*", "", d["Kernel Name"]).replace("<unnamed>::", "").replace("void ", ""), v, d["Grid Size"]))
    return out


def main():
    data = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for name, v, _ in data:
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print("launches %d  total %.1f us" % (len(data), total))
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        print("%-46s n=%4d %9.1f us %5.1f%%  avg %7.1f us" % (k[:46], cnt[k], v, 100 * v / total, v / cnt[k]))


if __name__ == "__main__":
    main()
