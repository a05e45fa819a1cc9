"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.md"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__occupancy_limit_registers", "CTAs/SM (reg limit)"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1 LSU wavefronts % of peak"),
        ("l1tex__t_sector_hit_rate.pct", "L1 sector hit %"), ("lts__t_sector_hit_rate.pct", "L2 sector hit %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "L1 global-RED sectors"),
        ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 RED sectors")]
print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"## {d['Kernel Name']}  (grid {d.get('Grid Size','?')}, block {d.get('Block Size','?')})\n")
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in d:
            print(f"| {label} (`{k}`) | {d[k]} {units[hdr.index(k)]} |")
    dr, dw = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
    print()

# optional: python tools/ncu_summary.py rep --traffic-json out.json  -> {"render_fwd": bytes, "render_bwd": bytes, ...}
if "--traffic-json" in sys.argv:
    import json, re
    out = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = re.sub(r"^void ", "", d["Kernel Name"]).split("<")[0].split("(")[0].replace("texgs_", "")
        def tobytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd = tobytes(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
        wr = tobytes(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
        out.setdefault(name, []).append(rd + wr)
    out = {k: int(sum(v) / len(v)) for k, v in out.items()}
    Path = __import__("pathlib").Path
    Path(sys.argv[sys.argv.index("--traffic-json") + 1]).write_text(json.dumps(out, indent=1) + "\n")
