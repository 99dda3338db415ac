"""Per-stage summary of a tools/step_trace.py dump (TRACE_ALL=1): time from one kernel's dependency wait to the next one's, and the
phases of the GEMMs (dep-ok -> first operand stage, k-loop, last stage -> accumulators ready, epilogue, end -> next dep-ok)."""
import sys,re,collections
for fn in sys.argv[1:]:
    rows=[]
    for l in open(fn):
        m=re.match(r"\s*([\d.]+) us\s+(\S+)\s+(\S+)\s+N=(\d+) K=(\d+)",l)
        if m: rows.append((float(m.group(1)),m.group(2),m.group(3),int(m.group(4)),int(m.group(5))))
    deps=[r for r in rows if r[1]=='dep-ok']
    def name(r):
        if r[2]=='attn': return 'attn'
        if r[2]=='rowln': return 'rowln%d'%r[4]
        return {3072:'qkv',1024:'proj1',2048:'proj',4096:'fc1/fc2'}.get(r[3],'other%d'%r[3])
    d=collections.defaultdict(list)
    for a,b in zip(deps,deps[1:]): d[name(a)].append(b[0]-a[0])
    print(fn, {k:'%.2f x%d'%(sum(v)/len(v),len(v)) for k,v in d.items()})
    q=[r[0] for r in deps if name(r)=='qkv']
    print('  block period', [round(b-a,1) for a,b in zip(q,q[1:])])
    # sub-phases of gemms: dep->stage0, stage0->stageL, stageL->acc, acc->end, end->next dep
    ph=collections.defaultdict(list)
    for i,r in enumerate(rows):
        if r[1]=='dep-ok' and r[2].startswith('gemm'):
            seq={}
            for s in rows[i+1:i+12]:
                if s[3]==r[3] and s[1] in('stage0','stageL','acc-ok','end') and s[1] not in seq: seq[s[1]]=s[0]
            nd=[s for s in rows[i+1:i+14] if s[1]=='dep-ok']
            if len(seq)==4 and nd:
                k=name(r)
                ph[k].append((seq['stage0']-r[0],seq['stageL']-seq['stage0'],seq['acc-ok']-seq['stageL'],seq['end']-seq['acc-ok'],nd[0][0]-seq['end']))
    for k,v in ph.items():
        n=len(v); print('  ',k,['%.2f'%(sum(x[i] for x in v)/n) for i in range(5)])
