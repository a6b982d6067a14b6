#!/usr/bin/env python
"""Hot SASS regions of an .ncu-rep: per-instruction executed counts, active lanes and stall samples.
    python tools/ncu_sass_hot.py rep [launch] [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')
blk = blocks[1 + launch]
lines = blk.split("\n")
print("kernel:", lines[0][:120])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ia, isrc, isamp, iex, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
ilong, iwait, ibr = hdr.index("stall_long_sb"), hdr.index("stall_wait"), hdr.index("stall_branch_resolving")
data = []
for r in rows[1:]:
    if len(r) <= ithr or not r[isamp]: continue
    try: data.append((int(r[isamp]), int(r[iex]), float(r[ithr] or 0), r[isrc], int(r[ilong] or 0), int(r[iwait] or 0), int(r[ibr] or 0)))
    except ValueError: pass
tot_s = sum(d[0] for d in data); tot_e = sum(d[1] for d in data)
print(f"instructions {len(data)}, samples {tot_s}, warp-instr executed {tot_e}, weighted avg threads {sum(d[1]*d[2] for d in data)/max(tot_e,1):.2f}")
print("top by samples:")
for i, d in sorted(enumerate(data), key=lambda x: -x[1][0])[:top]:
    print(f"  #{i:4d} samp {d[0]:6d} ({100*d[0]/tot_s:4.1f}%) exec {d[1]:9d} thr {d[2]:5.1f} long {d[4]:5d} wait {d[5]:5d} br {d[6]:4d}  {d[3][:70]}")
# histogram of executed counts by region (every 40 instructions)
print("regions (40-instr buckets): idx exec_sum(M) avg_thr samples%")
for b in range(0, len(data), 40):
    seg = data[b:b+40]; e = sum(x[1] for x in seg)
    print(f"  {b:4d} {e/1e6:9.1f} {sum(x[1]*x[2] for x in seg)/max(e,1):6.1f} {100*sum(x[0] for x in seg)/tot_s:5.1f}%")
