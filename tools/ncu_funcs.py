"""Aggregate an ncu SASS source page by enclosing function / kernel phase (innermost source line -> named range).
usage: ncu_funcs.py <src.csv> <nvdisasm.txt> <kernel-substring> <ranges: name:lo-hi,...>"""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1:4]
ranges = []
for part in sys.argv[4].split(','):
    nm, r = part.split(':'); lo, hi = r.split('-'); ranges.append((nm, int(lo), int(hi)))
line_of = {}; cur=None; infunc=False
for l in open(dis):
    if re.match(r'\s*\.section\s+\.text\.', l): infunc = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2)))
    m2 = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m2 and infunc: line_of[int(m2.group(1),16)] = cur
rows=list(csv.reader(open(src_csv)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='Address'); hdr=rows[hi]
ai,ii,si,ti=hdr.index('Address'),hdr.index('Instructions Executed'),hdr.index('# Samples'),hdr.index('Thread Instructions Executed')
base=int(rows[hi+1][ai],16)
agg=collections.defaultdict(lambda:[0,0,0]); tot=[0,0,0]
for r in rows[hi+1:]:
    if len(r)<=ii: continue
    f,n=line_of.get(int(r[ai],16)-base, ('?',0))
    name = f if f!='hcm_kernels.cu' else next((nm for nm,lo,h in ranges if lo<=n<=h), f'L{n}')
    v=(int(r[ii] or 0), int(r[si] or 0), int(r[ti] or 0))
    for k in range(3): agg[name][k]+=v[k]; tot[k]+=v[k]
print('total warp-instr', tot[0], 'samples', tot[1])
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:40]:
    print(f"{k:28s} instr {100*v[0]/tot[0]:5.1f}%  samples {100*v[1]/tot[1]:5.1f}%  active {v[2]/max(v[0],1):4.1f}")
