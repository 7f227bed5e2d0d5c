"""Where do the stall samples of a kernel sit, by SASS address?  Prints contiguous regions (split at big gaps in
sampled addresses) with their sample / instruction share and the opcodes that collect most samples.

    python tools/ncu_sass_regions.py <report.ncu-rep> <kernel-substring> [occurrence]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = [b for b in blocks if pat in b["name"]][occ]
h, data = b["rows"][0], b["rows"][1:]
ia, isrc, ism, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
ins = [(int(r[ia], 16), r[isrc], int(r[ism] or 0), int(r[iex] or 0)) for r in data]
base = ins[0][0]
tot_s = sum(x[2] for x in ins) or 1
tot_i = sum(x[3] for x in ins) or 1
print(f"{b['name'][:100]}\n samples {tot_s}  warp-instructions {tot_i}  sass {len(ins)}")
# regions: split where executed count changes by > 8x between neighbours (role / loop boundaries)
W = 64
for s in range(0, len(ins), W):
    chunk = ins[s:s + W]
    cs, ci = sum(x[2] for x in chunk), sum(x[3] for x in chunk)
    if cs / tot_s < 0.01 and ci / tot_i < 0.01:
        continue
    top = sorted(chunk, key=lambda x: -x[2])[:3]
    tops = "; ".join(f"{(x[0]-base)//16}:{x[1].split()[0] if x[1] else '?'}({100*x[2]/tot_s:.1f}%)" for x in top if x[2])
    print(f" sass[{s:5d}:{s+W:5d}] smp {100*cs/tot_s:5.1f}%  ins {100*ci/tot_i:5.1f}%  avg exec/inst {ci/len(chunk):10.0f}  {tops}")
