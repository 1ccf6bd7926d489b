"""Summarise `ncu --page raw --csv` dumps (one kernel launch each) into the few numbers the roofline needs:
duration, DRAM bytes, L2 / DRAM throughput, achieved occupancy, issue activity, top stall reasons.
usage: python scripts/ncu_summarize.py gpurun_out/r2_ncu_*_raw.csv > profiles/r2_ncu_summary.txt   (also writes JSON with --json PATH)"""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct"]
out = {}
args = [a for a in sys.argv[1:] if not a.startswith("--")]
jpath = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
if jpath in args:
    args.remove(jpath)
for path in args:
    rows = list(csv.reader(open(path)))
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    if hdr is None:
        print(path, ": no kernel rows"); continue
    units = rows[rows.index(hdr) + 1]
    data = [r for r in rows[rows.index(hdr) + 2:] if len(r) == len(hdr)]
    for r in data:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        if name.startswith("void "):
            name = name[5:]
        name = name.split("<")[0].split("::")[-1]   # the name bench.py's kernel timers use: no return type, namespace or template arguments
        ent = {}
        for k in WANT:
            if k in hdr:
                try:
                    ent[k] = float(r[hdr.index(k)].replace(",", ""))
                    ent[k + "__unit"] = units[hdr.index(k)]
                except ValueError:
                    pass
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h.split("issue_stalled_")[1].split("_per_issue")[0]] = float(r[i])
                except ValueError:
                    pass
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        def val(k, scale=1.0):
            return ent.get(k, float("nan")) * scale
        def to_bytes(k):
            u = ent.get(k + "__unit", "byte").lower()
            return val(k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        def to_ms(k):
            u = ent.get(k + "__unit", "ns").lower()
            return val(k) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(u, 1e-6)
        dur = to_ms("gpu__time_duration.sum")
        traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        out[name] = {"duration_ms": dur, "traffic_bytes_per_launch": traffic, "dram_read_bytes": to_bytes("dram__bytes_read.sum"),
                     "dram_write_bytes": to_bytes("dram__bytes_write.sum"), "dram_GBps": traffic / 1e9 / (dur / 1e3) if dur == dur and dur > 0 else None,
                     "dram_pct_of_peak": val("dram__throughput.avg.pct_of_peak_sustained_elapsed"), "l2_pct_of_peak": val("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                     "l2_hit_rate_pct": val("lts__t_sector_hit_rate.pct"), "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
                     "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warp_instructions": val("smsp__inst_executed.sum"),
                     "registers_per_thread": val("launch__registers_per_thread"), "grid": val("launch__grid_size"), "block": val("launch__block_size"),
                     "smem_dynamic": val("launch__shared_mem_per_block_dynamic"), "smem_static": val("launch__shared_mem_per_block_static"),
                     "top_stalls_cycles_per_issue": top, "source": path}
        e = out[name]
        print("%s\n  duration %.3f ms (cold-cache, under ncu)  DRAM read %.1f MB + write %.1f MB = %.1f MB  -> %.0f GB/s (%.1f %% of peak)  L2 %.1f %% of peak, hit rate %.1f %%"
              % (name, dur, e["dram_read_bytes"] / 1e6, e["dram_write_bytes"] / 1e6, traffic / 1e6, e["dram_GBps"] or 0, e["dram_pct_of_peak"], e["l2_pct_of_peak"], e["l2_hit_rate_pct"]))
        print("  grid %d x %d threads, %d regs/thread, smem %d + %d B; warps active %.1f %%, issue slots busy %.1f %%, %.3g warp instructions"
              % (e["grid"], e["block"], e["registers_per_thread"], e["smem_dynamic"], e["smem_static"], e["warps_active_pct"], e["issue_active_pct"], e["warp_instructions"]))
        print("  stalls (warp cycles stalled per issued instruction): " + ", ".join("%s %.2f" % kv for kv in top))
if jpath:
    json.dump(out, open(jpath, "w"), indent=1)
