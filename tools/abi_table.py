"""Emit the entry-point table of INTEGRATION.md from include/crfr.h: every exported function with the section it is declared
in and the reference interface (file:line) the nearest `ref:` comment above it cites.  Usage: python tools/abi_table.py"""
import re
import sys

src = open(sys.argv[1] if len(sys.argv) > 1 else "include/crfr.h").read()
section, ref = "", ""
rows = []
pos = 0
token = re.compile(r"/\*(.*?)\*/|\b(?:int|size_t|unsigned long long|const char\*)\s+(crfr_[a-z0-9_]+)\s*\(", re.S)
for m in token.finditer(src):
    if m.group(1) is not None:
        c = " ".join(l.strip().lstrip("*").strip() for l in m.group(1).splitlines()).strip()
        t = re.match(r"-{4,}\s*(.*?)\s*-{4,}$", c)
        if t:
            section, ref = t.group(1), ""
        r = re.search(r"ref:\s*(.*?)(?:\.\s|$)", c)
        if r:
            ref = r.group(1).strip().rstrip(".")
    else:
        rows.append((m.group(2), section, ref))
print("| entry point | group | reference interface it stands in for |")
print("|---|---|---|")
seen = set()
for name, sec, r in rows:
    if name in seen:
        continue
    seen.add(name)
    if not sec:
        r = ""                      # before the first group: error string, version, counters, switches
    r = re.split(r"\):\s|:\s\s|;\s", r)[0]
    if len(r) > 170:
        r = r[:170].rsplit(" ", 1)[0] + " ..."
    print("| `%s` | %s | %s |" % (name, sec or "library", r if r else "(plumbing: no counterpart in the reference)"))
