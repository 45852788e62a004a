#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per kernel, executed-instruction share and stall
samples by opcode class, plus the hottest SASS instructions.  (development aid)"""
import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
which=int(sys.argv[2]) if len(sys.argv)>2 else 0
ntop=int(sys.argv[3]) if len(sys.argv)>3 else 40
kern=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name':
        cur={'name':r[1],'rows':[]}; kern.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
k=kern[which]
hdr=k['rows'][0]; data=k['rows'][1:]
iS=hdr.index('# Samples'); iSrc=hdr.index('Source'); iE=hdr.index('Instructions Executed')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(r[iS] or 0) for r in data); totE=sum(int(r[iE] or 0) for r in data)
print(k['name'],'samples',tot,'warp-instr executed',totE)
byop=collections.Counter(); byopS=collections.Counter()
for r in data:
    op=r[iSrc].split()
    op=[o for o in op if not o.startswith('@')]
    o=op[0].split('.')[0] if op else '?'
    byop[o]+=int(r[iE] or 0); byopS[o]+=int(r[iS] or 0)
print('opcode: exec%  samples%')
for o,c in byop.most_common(28):
    print('  %-12s %5.1f%%  %5.1f%%'%(o,100*c/totE,100*byopS[o]/max(tot,1)))
agg=collections.Counter()
for r in data:
    for c in stall_cols: agg[hdr[c]]+=int(r[c] or 0)
print('stalls:',[(a,b) for a,b in agg.most_common(8)])
top=sorted(range(len(data)), key=lambda i:-int(data[i][iS] or 0))[:ntop]
for i in sorted(top):
    r=data[i]
    st={hdr[c][6:]:int(r[c] or 0) for c in stall_cols if int(r[c] or 0)>0}
    st=sorted(st.items(), key=lambda kv:-kv[1])[:2]
    print(i, r[iS], r[iE], r[iSrc][:80], st)
