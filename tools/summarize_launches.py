"""Summarise an ncu `--csv` launch list by kernel.  Usage: python tools/summarize_launches.py file.csv [top]

Works on `--metrics gpu__time_duration.sum` lists and on lists that also carry dram__bytes_read.sum / dram__bytes_write.sum
(one CSV row per launch and metric: rows are grouped by launch ID)."""
import collections
import csv
import re
import sys

_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    """-> list of (kernel name, microseconds, bytes read or None, bytes written or None, grid)"""
    rows = list(csv.reader(open(path)))
    hdr, launches = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        rec = launches.setdefault(d["ID"], {"name": name, "grid": d["Grid Size"]})
        v, u, m = float(d["Metric Value"].replace(",", "")), d["Metric Unit"], d["Metric Name"]
        if m.startswith("gpu__time_duration"):
            rec["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        elif m.startswith("dram__bytes_read"):
            rec["rd"] = v * _BYTES.get(u, 1.0)
        elif m.startswith("dram__bytes_write"):
            rec["wr"] = v * _BYTES.get(u, 1.0)
    return [(r["name"], r.get("us", 0.0), r.get("rd"), r.get("wr"), r["grid"]) for r in launches.values()]


def main():
    data = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    tot, cnt = collections.defaultdict(float), collections.Counter()
    rd, wr = collections.defaultdict(float), collections.defaultdict(float)
    has_dram = any(d[2] is not None for d in data)
    for name, v, r, w, _ in data:
        tot[name] += v
        cnt[name] += 1
        rd[name] += r or 0.0
        wr[name] += w or 0.0
    total = sum(tot.values())
    if has_dram:
        R, W = sum(rd.values()) / 1e9, sum(wr.values()) / 1e9
        print("launches %d  total %.1f us   DRAM read %.1f GB + write %.1f GB = %.1f GB" % (len(data), total, R, W, R + W))
    else:
        print("launches %d  total %.1f us" % (len(data), total))
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        line = "%-46s n=%4d %9.1f us %5.1f%%  avg %7.1f us" % (k[:46], cnt[k], v, 100 * v / total, v / cnt[k])
        if has_dram:
            line += "   dram %6.2f GB r + %6.2f GB w" % (rd[k] / 1e9, wr[k] / 1e9)
        print(line)


if __name__ == "__main__":
    main()
