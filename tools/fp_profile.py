#!/usr/bin/env python
"""Per-segment view of a fastparse ncu capture: tools/fp_profile.py <rep> -- instructions per round and stall samples
between the marker instructions of lz4_fastparse_unit (REDG = table insert, VOTE = selection, SHFL.UP = emission scan)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data)
mi = [i for i, r in enumerate(data) if "MATCH.ANY" in r[isrc]]
rounds = int(data[mi[0]][ia])
print("instructions", tot, "rounds", rounds, "per round %.1f" % (tot / rounds))
for nme in [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]:
    v = sum(int(r[hdr.index(nme)] or 0) for r in data)
    if v / ts > 0.01: print(nme, "%.1f%%" % (100 * v / ts), end="; ")
print()
first = [i for i, r in enumerate(data) if int(r[ia]) > rounds * 0.9][0]
marks = [first] + [i for i, r in enumerate(data) if int(r[ia]) > rounds * 0.9 and ("REDG" in r[isrc] or "VOTE.ANY" in r[isrc] or "SHFL.UP" in r[isrc] and "0x1," in r[isrc])] + [len(data)]
for a, b in zip(marks, marks[1:]):
    n = sum(int(r[ia]) for r in data[a:b]); s = sum(int(r[isamp]) for r in data[a:b])
    print("[%4d,%4d) %-44s per-round %6.1f  samples %5.1f%%" % (a, b, data[a][isrc].strip()[:44], n / rounds, 100 * s / ts))
if len(sys.argv) > 2:
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:int(sys.argv[2])]):
        print(i, "%5.2f%%" % (100 * int(data[i][isamp]) / ts), "x%.2f" % (int(data[i][ia]) / rounds), data[i][isrc][:90])
