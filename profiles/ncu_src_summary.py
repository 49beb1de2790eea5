#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instruction mix and stall samples by opcode class.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python ncu_src_summary.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
mix, samp = collections.Counter(), collections.Counter()
stalls = collections.Counter()
tot_ex = tot_s = 0
top = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break      # a second kernel matched the regex: summarise the first section only
    if len(r) <= iex or r[ia] == "Address": continue
    op = r[isrc].split()
    if not op: continue
    name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    name = name.split(".")[0]
    ex, s = int(r[iex] or 0), int(r[isamp] or 0)
    mix[name] += ex; samp[name] += s; tot_ex += ex; tot_s += s
    for i in stall_cols: stalls[hdr[i]] += int(r[i] or 0)
    top.append((s, r[ia], r[isrc][:70]))
print(f"total warp-instructions {tot_ex}, samples {tot_s}")
print("opcode      executed   %exec  %samples")
for k, v in mix.most_common(28):
    print(f"{k:10s} {v:10d} {100*v/tot_ex:6.2f} {100*samp[k]/max(tot_s,1):6.2f}")
print("stall reasons:", ", ".join(f"{k[6:]}={100*v/max(tot_s,1):.1f}%" for k, v in stalls.most_common(8)))
print("hottest instructions:")
for s, a, src in sorted(top, reverse=True)[:14]: print(f"  {100*s/max(tot_s,1):5.2f}%  {src}")
