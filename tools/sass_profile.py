"""Per-instruction listing of an `ncu --page source --csv --print-source sass` dump: executed count per
warp and stall samples per SASS instruction.  usage: sass_profile.py dump.csv [kernel-block-index] [n_warps]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append((r[1], cur))
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if cur is not None and len(r) > 10:
        cur.append(r)
name, b = blocks[blk]
iex, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
nw = float(sys.argv[3]) if len(sys.argv) > 3 else max(int(r[iex]) for r in b)
tot = sum(int(r[iex]) for r in b)
smp = sum(int(r[ismp]) for r in b)
print(name[:80])
print("warp-instructions %d  static %d  per warp %.1f  samples %d" % (tot, len(b), tot / nw, smp))
for i, r in enumerate(b):
    print("%4d %5.2f %5.1f%%  %s" % (i, int(r[iex]) / nw, 100.0 * int(r[ismp]) / max(smp, 1), r[isrc].strip()))
