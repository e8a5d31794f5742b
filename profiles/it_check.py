import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from mutation_simulator_b200.engine import Engine, BUF_FASTA
from mutation_simulator_b200.records import REC_DTYPE, K_RAW, T_IT
L = bench.GRCH38; names=[n.encode() for n in bench.GRCH38_NAMES]
eng = Engine(0)
eng.synth_genome(1, L, [60]*24, names, names, 0.03, 10000)
rng = np.random.default_rng(0)
perm = rng.permutation(24); pairs=[(int(perm[2*i]), int(perm[2*i+1])) for i in range(12)]
for rate in (1e-7, 1e-4):
    counts=[int((L[a]+L[b]-4)/2*rate) for a,b in pairs]
    bpa,bpb = eng.it_breakpoints(5,[p[0] for p in pairs],[p[1] for p in pairs],counts)
    goff=np.concatenate(([0],np.cumsum(L)))
    parts=[];o=0
    bps={}
    for (a,b),n in zip(pairs,counts):
        bps[a]=(bpa[o:o+n],bpb[o:o+n],b); bps[b]=(bpb[o:o+n],bpa[o:o+n],a); o+=n
    for c in sorted(bps):
        s,pp,p=bps[c]
        A=np.concatenate(([0],s.astype(np.int64),[L[c]])); B=np.concatenate(([0],pp.astype(np.int64),[L[p]]))
        odd=np.arange(1,len(A)-1,2); r=np.zeros(len(odd),REC_DTYPE)
        r['pos']=A[odd]; r['cons']=A[odd+1]-A[odd]; r['prod']=B[odd+1]-B[odd]; r['src']=goff[p]+B[odd]; r['kind']=K_RAW; r['type']=T_IT; r['contig']=c
        parts.append(r)
    recs=np.concatenate(parts)
    eng.load_records(recs)
    eng.apply(); eng.apply()
    st=eng.stats()
    print('rate',rate,'breakpoints',sum(counts),'records',len(recs),'fasta',st['fasta_bytes'],{k:round(v,3) for k,v in st['stage_ms'].items() if v>0.001})
    img=eng.download(BUF_FASTA)
    assert st['fasta_bytes']>3.1e9
eng.close()
