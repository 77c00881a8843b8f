"""Aggregate the warp-stall samples of an ncu report per SOURCE LINE without a GPU: ncu's SASS page (samples per instruction) is
joined with nvdisasm -g (source line per instruction) by instruction index.
    python tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-substring> [top]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], rows=[]); blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)
blk = next(b for b in blocks if kern in b["name"])
h = blk["hdr"]; ci = h.index("Warp Stall Sampling (All Samples)")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
lines, infn, cur_line = [], False, None
mangled_hint = kern
for l in dis.splitlines():
    if l.startswith(".text."):
        infn = mangled_hint.replace("<", "").split("(")[0].split("::")[-1].split("<")[0] in l and ("ILi" + kern.split("<")[1].split(">")[0].replace("(int)", "") + "E" in l if "<" in kern else True)
        cur_line = None
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur_line)
n = min(len(lines), len(blk["rows"]))
print(f"{blk['name'][:80]}: {len(blk['rows'])} SASS instructions in the report, {len(lines)} in the object")
agg = collections.defaultdict(lambda: [0, collections.Counter()])
tot = 0
for i in range(n):
    r = blk["rows"][i]; s = int(r[ci] or 0); tot += s
    a = agg[lines[i]]; a[0] += s
    for j in stall:
        v = int(r[j] or 0)
        if v: a[1][h[j]] += v
src_cache = {}
def src(fl):
    if fl is None: return ""
    f, ln = fl
    if f not in src_cache:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
        src_cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
    L = src_cache[f]
    return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
print("total samples", tot)
for fl, (s, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{s:6d} {100*s/max(tot,1):5.1f}%  {fl[0] if fl else '?'}:{fl[1] if fl else 0:<5d} {dict(c.most_common(2))}  | {src(fl)}")
