"""Attribute ncu warp-stall samples (SASS page) to CUDA source lines via nvdisasm line info.

usage: ncu_lines.py report.ncu-rep lib.so kernel_substring source_file.cu [top_n]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kname, srcfile = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines_of = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    if kname not in out:
        continue
    in_k, cur, ctx = False, None, None
    for ln in out.splitlines():
        if ln.startswith(".text."):
            in_k = kname in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            base = os.path.basename(m.group(1))
            if base == os.path.basename(srcfile):
                ctx = (base, int(m.group(2))); cur = ctx
            elif os.path.exists(os.path.join(os.path.dirname(srcfile), base)):
                cur = (base, int(m.group(2)))  # our own header next to the source file
            else:
                cur = ctx  # toolkit header code: charge the enclosing line of our file
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            lines_of.append((cur, ln.split("*/", 1)[1].strip()))
    if lines_of:
        break
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci = {k: i for i, k in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) > 5 and r[0].startswith("0x")]
print("sass instrs: ncu %d, nvdisasm %d" % (len(body), len(lines_of)))
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
for i, r in enumerate(body):
    line = lines_of[i][0] if i < len(lines_of) else None
    s = int(r[ci["# Samples"]] or 0)
    tot += s
    a = agg[line]
    a["samples"] += s
    a["inst"] += int(r[ci["Instructions Executed"]] or 0)
    for k in stalls:
        v = r[ci[k]]
        if v and v != "0":
            a[k] += int(v)
srcs = {}
def text_of(key):
    if not key:
        return "?"
    base, line = key
    if base not in srcs:
        srcs[base] = open(os.path.join(os.path.dirname(srcfile), base)).read().splitlines()
    return srcs[base][line - 1].strip()[:80] if line <= len(srcs[base]) else "?"
print("total samples", tot)
glob = collections.Counter()
for a in agg.values():
    for k in stalls:
        glob[k] += a[k]
print("stall mix:", [(k, "%.1f%%" % (100.0 * v / tot)) for k, v in glob.most_common(8)])
for line, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    top = sorted(((k, a[k]) for k in stalls if a[k]), key=lambda x: -x[1])[:3]
    where = "%s:%d" % (line[0].replace("chomp_", "").split(".")[0], line[1]) if line else "?"
    print("%-12s %5.1f%% inst=%-10d %-80s %s" % (where, 100.0 * a["samples"] / tot, a["inst"], text_of(line), [(k[6:], v) for k, v in top]))
