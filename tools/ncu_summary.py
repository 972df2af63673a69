"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).  Usage:
   python tools/ncu_summary.py <tag> <launches.csv> <prof.ncu-rep> <kernel-regex>"""
import collections, csv, json, os, re, subprocess, sys
tag, launches, rep, kre = sys.argv[1:5]
out = [f"# ncu summary {tag}", ""]
if os.path.exists(launches):
    rows = list(csv.reader(open(launches))); hdr = None; agg = collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try: v = float(d["Metric Value"].replace(",", ""))
            except ValueError: continue
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1)
            a = agg.setdefault(d["Kernel Name"][:90], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out += [f"## launch list ({os.path.basename(launches)}; `ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache serialised: compare shares)", "",
            "| total us | launches | share | kernel |", "|---:|---:|---:|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
        out.append(f"| {t:.1f} | {n} | {100*t/tot:.1f}% | `{k}` |")
    out.append("")
if os.path.exists(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines())); hdr = rows[0]; units = dict(zip(hdr, rows[1]))
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "sm__inst_executed.avg.per_cycle_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__warps_eligible.avg.per_cycle_active", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum"]
    out += [f"## `ncu --set full --clock-control none` ({os.path.basename(rep)})", ""]
    traffic = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if not re.search(kre, d.get("Kernel Name", "")): continue
        out.append(f"### {d['Kernel Name']}  grid {d.get('launch__grid_size')} x block {d.get('launch__block_size')}")
        out += ["", "| metric | value | unit |", "|---|---:|---|"]
        for w in want:
            if w in d and d[w] != "": out.append(f"| {w} | {d[w]} | {units.get(w,'')} |")
        def tob(k):
            v = float(d[k].replace(",", "")); u = units[k]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        try:
            traffic = tob("dram__bytes_read.sum") + tob("dram__bytes_write.sum")
            out.append(f"| dram bytes per launch (read+write) | {traffic:.0f} | byte |")
        except Exception as e:
            out.append(f"| traffic | n/a ({e}) | |")
        out.append("")
    if traffic:
        json.dump({"dram_bytes_per_launch": traffic, "source": os.path.basename(rep), "kernel": kre},
                  open(os.path.join("profiles", f"traffic_{tag}.json"), "w"))
open(os.path.join("profiles", f"{tag}.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
