"""Attribute ncu warp-stall samples to CUDA source lines without a GUI: ncu's `--page source --csv` gives samples per
SASS instruction; `nvdisasm -g` of the same kernel gives the source line of every SASS instruction; both list the
instructions in the same order.  Usage:
    python tools/ncu_source_lines.py <report.ncu-rep> <cubin> <mangled-or-substring-of-kernel-name> [top_n]
(cubins: cuobjdump -xelf all elliot_b200/csrc/libelliot_b200.so)"""
import collections, csv, re, subprocess, sys

rep, cubin, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
hi = [i for i, l in enumerate(txt) if l.startswith('"Address"')][0]
rows = list(csv.reader(txt[hi:])); hdr, data = rows[0], rows[1:]
ia = hdr.index("Warp Stall Sampling (All Samples)")
samples = [int(r[ia]) if r[ia].isdigit() else 0 for r in data]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ridx = {h: hdr.index(h) for h in reasons}
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# isolate the function
blocks = re.split(r"\n\s*\.section\s+\.text\.", dis)
blk = next((b for b in blocks if kname in b.split("\n", 1)[0]), None)
if blk is None:
    sys.exit(f"kernel {kname} not found in {cubin}")
line_of, cur = [], None
for l in blk.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4}\*/\s+", l):
        line_of.append(cur)
if len(line_of) != len(samples):
    sys.exit(f"instruction count mismatch: ncu {len(samples)} vs nvdisasm {len(line_of)} (different build?)")
agg = collections.Counter(); why = collections.defaultdict(collections.Counter)
for k, (s, ln) in enumerate(zip(samples, line_of)):
    agg[ln] += s
    for h in reasons:
        v = data[k][ridx[h]]
        if v.isdigit() and int(v):
            why[ln][h[6:]] += int(v)
tot = sum(samples)
src_cache = {}
print(f"total samples {tot}; top {topn} source lines")
for (f, ln), s in agg.most_common(topn) if None not in agg else [(k, v) for k, v in agg.most_common(topn) if k]:
    if f not in src_cache:
        try: src_cache[f] = open(f"elliot_b200/csrc/{f}").read().splitlines()
        except OSError: src_cache[f] = []
    code = src_cache[f][ln - 1].strip()[:90] if 0 < ln <= len(src_cache[f]) else ""
    top_why = ", ".join(f"{k} {100 * v // max(s, 1)}%" for k, v in why[(f, ln)].most_common(2))
    print(f"{100 * s / tot:5.1f}%  {f}:{ln:<4d} [{top_why}]  {code}")
