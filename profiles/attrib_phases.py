import csv, re, sys, collections
dis, sasscsv, func = sys.argv[1:4]
lines=open(dis).read().split('\n')
cur_line=None; cur_file=None; ins=[]; infunc=False
for l in lines:
    m=re.match(r'\s*\.text\.(\S+):',l)
    if m: infunc = func in m.group(1); continue
    if not infunc: continue
    m=re.match(r'\s*//## File "(.*)", line (\d+)(.*)',l)
    if m: cur_file=m.group(1).split('/')[-1]; cur_line=int(m.group(2)); continue
    if re.match(r'\s*/\*[0-9a-f]{4,6}\*/',l): ins.append((cur_file,cur_line,l.strip()))
rows=list(csv.reader(open(sasscsv))); hdr=rows[1]; data=[r for r in rows[2:] if len(r)>=10]
ixs=hdr.index('# Samples'); ixi=hdr.index('Instructions Executed')
# phases by k3_windows.cu line ranges of the *call site*: approximate using nearest preceding k3_windows.cu line
phase_bounds=[(0,'prolog'),(268,'load+nan'),(278,'zscore'),(305,'std'),(311,'qrs-fast'),(367,'qrs-exact'),(400,'runs'),(432,'minlen'),(447,'peaks-per-run'),(484,'thread0-stats'),(536,'peaks-out+median')]
def phase_of(line):
    p='prolog'
    for b,n in phase_bounds:
        if line>=b: p=n
    return p
agg=collections.OrderedDict((n,[0,0]) for _,n in phase_bounds)
last_k3=0
opc=collections.Counter()
for k in range(min(len(ins),len(data))):
    f,ln,txt=ins[k]
    if f=='k3_windows.cu' and ln>=245: last_k3=ln
    ph=phase_of(last_k3)
    agg[ph][0]+=int(data[k][ixs]); agg[ph][1]+=int(data[k][ixi])
    op=re.sub(r'/\*[0-9a-f]+\*/\s*','',txt).split()[0] if txt else ''
    if op.startswith('@'): op=re.sub(r'/\*[0-9a-f]+\*/\s*','',txt).split()[1]
    opc[op.split('.')[0]]+=int(data[k][ixi])
ts=sum(v[0] for v in agg.values()); ti=sum(v[1] for v in agg.values())
for n,v in agg.items(): print(f"{n:20s} samples {v[0]/ts*100:5.1f}%  inst {v[1]/ti*100:5.1f}%  ({v[1]/11456/1e3:.1f}k warp-inst/window)")
print({k:round(v/ti*100,1) for k,v in opc.most_common(25)})
