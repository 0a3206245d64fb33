"""Per-source-line attribution of an ncu capture: aligns the SASS page of the report (ncu -i X.ncu-rep --page source
--csv --print-source sass) with nvdisasm -g -c of the cubin (same build), then sums samples, executed
instructions and the main stall reasons per source line.  The kernel's non-inlined device functions follow it in
both listings; functions are taken in file order from the kernel's label until the instruction counts agree.
usage: attrib_lines.py <nvdisasm.txt> <sass.csv> <mangled-name-fragment> [top_n] [by=samples|inst]"""
import csv, re, sys, collections
dis, sasscsv, func = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by = 1 if (len(sys.argv) > 5 and sys.argv[5] == "inst") else 0
rows = list(csv.reader(open(sasscsv))); hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= 46]
funcs = []; cur = None; name = None
for l in open(dis).read().split('\n'):
    m = re.match(r'^(\S+):\s*$', l)
    if m and not m.group(1).startswith('.L'):
        name = m.group(1); funcs.append((name, []))
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,6}\*/', l) and funcs: funcs[-1][1].append((cur, name))
start = [k for k, (n, v) in enumerate(funcs) if func in n and not n.startswith('$')][0]
ins = []
for n, v in funcs[start:]:
    if len(ins) >= len(data): break
    ins += v
print("sass rows", len(data), "disasm instructions", len(ins))
assert len(data) == len(ins), "report and cubin are different builds"
cols = ['# Samples', 'Instructions Executed', 'stall_barrier', 'stall_no_inst', 'stall_wait', 'stall_short_sb',
        'stall_long_sb', 'stall_math', 'stall_branch_resolving', 'stall_mio']
ix = [hdr.index(c) for c in cols]
agg = collections.defaultdict(lambda: [0] * len(cols)); fagg = collections.defaultdict(lambda: [0, 0, 0])
for (key, fn), r in zip(ins, data):
    a = agg[key]
    for k, i in enumerate(ix): a[k] += int(r[i])
    f = fagg[fn]; f[0] += int(r[ix[0]]); f[1] += int(r[ix[1]]); f[2] += 1
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print("per function: samples%  inst%  static instructions")
for fn, v in sorted(fagg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {v[0] / ts * 100:5.1f}% {v[1] / ti * 100:5.1f}% {v[2]:6d}  {fn[-44:]}")
src = {}
def text(f, ln):
    if f not in src:
        try: src[f] = open('/root/repo/ae_waveforms_b200/csrc/' + f).read().split('\n')
        except Exception: src[f] = []
    return src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ''
print("line  samples%  inst%  [barrier no_inst wait short_sb long_sb math branch mio as % of the line's samples]")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
    rs = ' '.join(f"{100 * x / max(v[0], 1):3.0f}" for x in v[2:])
    print(f"{f}:{ln:<4} {v[0] / ts * 100:5.1f}% {v[1] / ti * 100:5.1f}%  [{rs}]  {text(f, ln)}")
