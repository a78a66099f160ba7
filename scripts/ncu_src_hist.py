"""Dynamic opcode histogram of an `ncu --page source --csv` dump, per 6 trellis stages of one warp.
usage: ncu_src_hist.py src.csv <warps> <stages per warp>"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1] if rows[0][0] == "Kernel Name" else rows[0]
data = rows[2:] if rows[0][0] == "Kernel Name" else rows[1:]
warps, stages = float(sys.argv[2]), float(sys.argv[3])
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
per6 = warps * stages / 6
tot = 0
byop, samp = collections.Counter(), collections.Counter()
for r in data:
    n = int(r[iE]); tot += n
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)?)', r[iS].strip())
    op = m.group(2) if m else r[iS][:10]
    if not op.startswith(("SHFL", "LDS", "STS", "IMAD.MOV", "IMAD.IADD", "VIMNMX", "HMNMX2", "HSETP2", "HADD2", "HFMA2", "LDG", "STG")):
        op = op.split(".")[0]
    if m and m.group(1) and op == "IMAD":
        op = "@P IMAD"
    byop[op] += n; samp[op] += int(r[iW])
print("total warp instructions %d = %.1f per 6 stages per warp" % (tot, tot / per6))
for op, n in byop.most_common(40):
    print("%-14s %8.2f   stall samples %6d" % (op, n / per6, samp[op]))
