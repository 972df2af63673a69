"""Top CUDA source lines by warp-stall samples straight from an ncu report captured with --import-source on:
    python tools/ncu_cuda_lines.py <report.ncu-rep> [top_n]
(uses `ncu --page source --csv --print-source cuda,sass`: rows with a line number carry the line's aggregated samples)."""
import csv, subprocess, sys

rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout.splitlines()
rows, cur_file, hdr = [], None, None
for r in csv.reader(txt):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        s = d.get("Warp Stall Sampling (All Samples)", "0")
        rows.append((int(s) if s.isdigit() else 0, cur_file, int(r[0]), r[1].strip()[:100], d))
tot = sum(x[0] for x in rows) or 1
print(f"total samples {tot}")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h] if hdr else []
for s, f, ln, code, d in sorted(rows, key=lambda x: -x[0])[:topn]:
    why = sorted(((int(d[h]) if d[h].isdigit() else 0, h[6:]) for h in reasons), reverse=True)[:2]
    print(f"{100 * s / tot:5.1f}%  {f}:{ln:<4d} [{', '.join(f'{n} {100 * v // max(s, 1)}%' for v, n in why)}]  {code}")
