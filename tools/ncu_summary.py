#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) per kernel: key metrics + top stall reasons.  Runs on the CPU box."""
import csv, subprocess, sys, io
reps = [a for a in sys.argv[1:] if a.endswith((".csv", ".ncu-rep"))]
rows = []
for rep in reps:                                   # several captures (one per pairing layout) are concatenated
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    rows = rr if not rows else rows + rr[2:]
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct', 'l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:60])
    for h, u, v in zip(hdr, rows[1], r):
        if h in want:
            print('   %-75s %s %s' % (h, v, u))
    items = []
    for h, v in zip(hdr, r):
        if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            try:
                items.append((float(v), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print('   stalls (cycles per issue):', ', '.join('%s %.2f' % (h, v) for v, h in sorted(items, reverse=True)[:8]))

# --traffic-json OUT BATCH: dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel (largest launch of each
# kernel name: the 3-thread key-side gather of the setup is not the one bench.py reports), read by bench.py's roofline.traffic
if "--traffic-json" in sys.argv:
    import json, re
    out, batch = sys.argv[sys.argv.index("--traffic-json") + 1], int(sys.argv[sys.argv.index("--traffic-json") + 2])
    UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    units = dict(zip(hdr, rows[1]))
    kern = {}
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        name = re.sub(r"<.*", "", rec["Kernel Name"].split("(")[0].replace("void ", "").replace("rb::", "")).strip()
        try:
            b = sum(float(rec[m]) * UNIT[units[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            ms = float(rec["gpu__time_duration.sum"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units["gpu__time_duration.sum"]]
        except (KeyError, ValueError):
            continue
        if name not in kern or b > kern[name]["dram_bytes_per_launch"]:
            kern[name] = {"dram_bytes_per_launch": b, "ncu_ms": ms, "registers": rec.get("launch__registers_per_thread"),
                          "grid": rec.get("launch__grid_size"), "block": rec.get("launch__block_size")}
    json.dump({"source": "ncu --set full --clock-control none over `python bench.py --steps 1 --warmup 3` (tools/gpu_profile_round.sh)",
               "batch": batch, "kernels": kern}, open(out, "w"), indent=1)
