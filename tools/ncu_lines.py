"""Aggregate an ncu SASS source page (--page source --csv) by CUDA source line, using nvdisasm -g line info.
usage: ncu_lines.py <src.csv> <nvdisasm.txt> <kernel-substring> [top]"""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: find function, then track "//## File "...", line N" and instruction offsets /*0000*/
line_of = {}
cur = None; infunc = False
for l in open(dis):
    if l.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.', l):
        infunc = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)), 'inlined' in m.group(3))
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m and infunc: line_of[int(m.group(1),16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = next(i for i,r in enumerate(rows) if r and r[0]=='Address')
hdr = rows[hi]
ai, ii, si = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ti = hdr.index('Thread Instructions Executed')
base = int(rows[hi+1][ai],16)
agg = collections.defaultdict(lambda:[0,0,0])
tot=[0,0,0]
for r in rows[hi+1:]:
    if len(r) <= ii: continue
    off = int(r[ai],16)-base
    key = line_of.get(off, ('?',0,False))[:2]
    n, s, t = int(r[ii] or 0), int(r[si] or 0), int(r[ti] or 0)
    agg[key][0]+=n; agg[key][1]+=s; agg[key][2]+=t
    tot[0]+=n; tot[1]+=s; tot[2]+=t
print('total warp-instr', tot[0], 'samples', tot[1], 'thread-instr', tot[2], 'avg active', tot[2]/max(tot[0],1))
srcs={}
def srcline(f,n):
    import glob
    if f not in srcs:
        c=glob.glob('/root/repo/hypercraft_b200/csrc/'+f)
        srcs[f]=open(c[0]).read().split('\n') if c else []
    return srcs[f][n-1].strip()[:110] if 0<n<=len(srcs[f]) else ''
for key,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
    print(f"{key[0]}:{key[1]:4d}  instr {100*v[0]/tot[0]:5.1f}%  samples {100*v[1]/max(tot[1],1):5.1f}%  act {v[2]/max(v[0],1):4.1f} | {srcline(*key)}")
