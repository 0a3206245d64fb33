import csv, re, sys, collections
dis, sasscsv, func, srcfile = sys.argv[1:5]
# parse nvdisasm: per function list of (line, file) per instruction in order
lines=open(dis).read().split('\n')
cur=None; cur_line=None; cur_file=None; ins=[]
infunc=False
for l in lines:
    m=re.match(r'\s*\.text\.(\S+):',l)
    if m:
        infunc = func in m.group(1); continue
    if not infunc: continue
    m=re.match(r'\s*//## File "(.*)", line (\d+)',l)
    if m: cur_file=m.group(1); cur_line=int(m.group(2)); continue
    if re.match(r'\s*/\*[0-9a-f]{4,6}\*/',l):
        ins.append((cur_file,cur_line,l.strip()))
rows=list(csv.reader(open(sasscsv)))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)>=10]
ixs=hdr.index('# Samples'); ixi=hdr.index('Instructions Executed')
print("sass rows",len(data),"disasm ins",len(ins))
n=min(len(data),len(ins))
agg=collections.defaultdict(lambda:[0,0])
for k in range(n):
    f,ln,_=ins[k]; key=(f.split('/')[-1] if f else '?',ln)
    agg[key][0]+=int(data[k][ixs]); agg[key][1]+=int(data[k][ixi])
ts=sum(v[0] for v in agg.values()); ti=sum(v[1] for v in agg.values())
src={}
for fn in set(k[0] for k in agg):
    try: src[fn]=open('/root/repo/ae_waveforms_b200/csrc/'+fn).read().split('\n')
    except Exception: src[fn]=[]
for (f,ln),v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[5]) if len(sys.argv)>5 else 40]:
    text = src.get(f,[])[ln-1].strip()[:100] if ln and ln-1 < len(src.get(f,[])) else ''
    print(f"{f}:{ln:<4} smp={v[0]/ts*100:5.1f}% inst={v[1]/ti*100:5.1f}%  {text}")
