"""Summarise `ncu --page raw --csv` exports into the text files under profiles/ and the per-kernel DRAM traffic
table bench.py reads (profiles/r2_traffic.json).

    python profiles/ncu_summary.py <name>=<raw.csv>:<units per launch>:<unit>[:mean|sum|each] ...
    e.g. k2_pass=gpurun_out/r2base/k2.raw.csv:115200000:sample:mean   (mean over the kernels of the file = per launch)
         k1=gpurun_out/r2base/k1.raw.csv:180000000:sample:sum         (the four launches of one K1 stage added up)
         k3=gpurun_out/r2k3/k3.raw.csv:17900:window:each               (one entry per kernel name)
"""
import csv, json, os, re, sys
HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__icc_request_hit_rate.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}
traffic_path = os.path.join(HERE, 'r2_traffic.json')
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for arg in sys.argv[1:]:
    name, rest = arg.split('=', 1)
    parts = rest.split(':')
    path, units, unit = parts[:3]
    how = parts[3] if len(parts) > 3 else "mean"
    per_kernel = []
    rows = list(csv.reader(open(path)))
    hdr, unitrow = rows[0], rows[1]
    out = [f"ncu --set full --clock-control none, raw page of {os.path.basename(path)} ({units} {unit}s per launch)"]
    seen = {}
    for r in rows[2:]:
        kn = re.sub(r'\(.*', '', r[hdr.index('Kernel Name')]).replace('<unnamed>::', '').strip()
        if kn in seen:
            continue                                   # first launch of every kernel
        seen[kn] = True
        out.append(f"==== {kn}")
        vals = {}
        for i, h in enumerate(hdr):
            if h in KEEP or ('pcsamp_warps_issue_stalled' in h and 'not_issued' not in h):
                out.append(f"  {h:82s} {r[i]} {unitrow[i]}")
            if h in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                vals[h] = float(r[i].replace(',', '')) * UNIT.get(unitrow[i], 1.0)
        per_kernel.append((kn.replace('void ', '').split('<')[0], vals['dram__bytes_read.sum'], vals['dram__bytes_write.sum']))
    def entry(rd, wr):
        return {"dram_bytes_per_unit": (rd + wr) / float(units), "unit": unit, "source": f"profiles/r2_ncu_{name}.txt",
                "dram_bytes_read": rd, "dram_bytes_write": wr, "units_per_launch": float(units)}
    if how == "each":
        for k, rd, wr in per_kernel:
            traffic[k] = entry(rd, wr)
    else:
        d = len(per_kernel) if how == "mean" else 1
        traffic[name] = entry(sum(p[1] for p in per_kernel) / d, sum(p[2] for p in per_kernel) / d)
        traffic[name]["kernels"] = [p[0] for p in per_kernel]
    open(os.path.join(HERE, f"r2_ncu_{name}.txt"), 'w').write('\n'.join(out) + '\n')
    print(name, list(seen))
json.dump(traffic, open(traffic_path, 'w'), indent=1)
