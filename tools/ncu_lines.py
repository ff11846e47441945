"""Attribute an ncu report's executed instructions / stall samples to CUDA source lines.
usage: python tools/ncu_lines.py <report.ncu-rep> <lib.so> <mangled kernel substring> [top]"""
import collections, csv, io, re, subprocess, sys, os, tempfile
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
cur, seq = None, []
for l in sass[start + 1:]:
    if l.startswith("//--------------------- .") and seq:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(2).strip(), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
prof = [(r[1].strip(), int(r[ci]), int(r[si]) if r[si].isdigit() else 0) for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
assert len(prof) == len(seq), (len(prof), len(seq))
byline, bysamp, byop = collections.Counter(), collections.Counter(), collections.Counter()
for (txt, line), (ptxt, n, smp) in zip(seq, prof):
    byline[line] += n
    bysamp[line] += smp
    op = [t for t in ptxt.split() if not t.startswith("@")][0].split(".")[0]
    byop[op] += n
tot, ts = sum(byline.values()), sum(bysamp.values())
print("total warp-instructions", tot, "samples", ts, "SASS length", len(seq))
print("by opcode:", ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in byop.most_common(14)))
srcdir = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc")
cache = {}
order = bysamp.most_common(top) if os.environ.get("BY_SAMPLES") else byline.most_common(top)
for k, _v in order:
    v = byline[k]
    f, ln = k if k else ("?", 0)
    if f not in cache:
        pth = os.path.join(srcdir, f)
        cache[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
    text = cache[f][ln - 1].strip()[:100] if 0 < ln <= len(cache[f]) else ""
    print(f"{v:10d} {100*v/tot:5.1f}% stall-samples {100*bysamp[k]/max(ts,1):5.1f}%  {f}:{ln}: {text}")
