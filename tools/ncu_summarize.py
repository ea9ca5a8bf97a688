"""Text summary of an .ncu-rep (key raw metrics per kernel + top stall instructions).
Usage: python tools/ncu_summarize.py gpurun_out/prof.ncu-rep > profiles/xxx.txt   (needs the ncu CLI)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}")
for r in rows[2:]:
    print(f"\n## kernel: {r[ix['Kernel Name']]}")
    for w in WANT:
        if w in ix:
            print(f"{w:75s} {r[ix[w]]:>16s} {units[ix[w]]}")
    try:
        rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[ix["dram__bytes_read.sum"]], 1.0)
        t = float(r[ix["gpu__time_duration.sum"]]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(units[ix["gpu__time_duration.sum"]], 1e-6)
        print(f"{'derived: DRAM traffic / time':75s} {(rd + wr) * scale / t / 1e9:16.1f} GB/s (under ncu, cold clocks)")
    except Exception:
        pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if not b["rows"]:
        continue
    h = b["rows"][0]
    jx = {c: i for i, c in enumerate(h)}
    data = [r for r in b["rows"][1:] if len(r) >= len(h)]
    stall = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[jx["# Samples"]] or 0) for r in data) or 1
    agg = {c[6:]: sum(int(r[jx[c]] or 0) for r in data) for c in stall}
    print(f"\n## stall samples: {b['name'][:90]}")
    print("   ", {k: f"{100.0 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 > tot})
    for r in sorted(data, key=lambda r: -int(r[jx["# Samples"]] or 0))[:8]:
        st = {c[6:]: int(r[jx[c]] or 0) for c in stall if int(r[jx[c]] or 0) * 20 > int(r[jx["# Samples"]] or 1)}
        print(f"    {100.0 * int(r[jx['# Samples']] or 0) / tot:5.1f}%  {r[jx['Source']][:60]:60s} {st}")
