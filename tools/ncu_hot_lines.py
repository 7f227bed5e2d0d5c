"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

    python tools/ncu_hot_lines.py <report.ncu-rep> <kernel-name-substring> [top N]

Joins `ncu --page source --csv` (SASS view: executed instructions + stall samples per instruction) with
`nvdisasm -g` line info of the matching cubin extracted from libcimq.so (built with -lineinfo)."""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile

rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the report may hold several kernels: split on "Kernel Name" header rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
blk = [b for b in blocks if pat in b["name"]][0]
hdr, data = blk["rows"][0], blk["rows"][1:]
ia, iex, ism, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
base = min(int(r[ia], 16) for r in data)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "cim_quantization_b200", "libcimq.so")], cwd=tmp,
               capture_output=True)
linemap = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    # find the function whose demangled-ish name matches
    sections = re.split(r"\n//-+ \.text\.", dis)
    for sec in sections[1:]:
        mangled = sec.split(" ", 1)[0]
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout
        if pat in dem and blk["name"].split("(")[0].replace("void ", "").strip()[:40] in dem.replace("(int)", ""):
            pass
        if pat in dem:
            m, fileline = {}, ("?", 0)
            for ln in sec.splitlines():
                mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if mm:
                    fileline = (os.path.basename(mm.group(1)), int(mm.group(2)))
                    continue
                mi = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
                if mi:
                    m[int(mi.group(1), 16)] = fileline
            if len(m) == len(data) or linemap is None:
                linemap = m
            if len(m) == len(data):
                break
agg = collections.defaultdict(lambda: [0, 0])
tot_ex = tot_sm = 0
for r in data:
    off = int(r[ia], 16) - base
    fl = linemap.get(off, ("?", 0))
    ex, sm = int(r[iex]), int(r[ism])
    agg[fl][0] += ex
    agg[fl][1] += sm
    tot_ex += ex
    tot_sm += sm
src_cache = {}
def src(fl):
    f, l = fl
    for d in ("cim_quantization_b200/csrc",):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:90] if 0 < l <= len(src_cache[p]) else ""
    return ""
print(f"kernel: {blk['name'][:100]}\ninstructions executed {tot_ex}, stall samples {tot_sm}, sass {len(data)} (mapped {len(linemap)})")
print("--- by stall samples")
for fl, (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{100*sm/max(tot_sm,1):5.1f}% smp {100*ex/max(tot_ex,1):5.1f}% ins  {fl[0]}:{fl[1]:<5d} {src(fl)}")
print("--- by instructions executed")
for fl, (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*ex/max(tot_ex,1):5.1f}% ins {100*sm/max(tot_sm,1):5.1f}% smp  {fl[0]}:{fl[1]:<5d} {src(fl)}")
