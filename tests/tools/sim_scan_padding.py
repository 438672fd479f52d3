"""Development tool (test infrastructure: it drives the CPU oracle, so it lives under tests/).

How many MMAs would the scan execute under different blockings?  The oracle's `_grid_argmax` is wrapped to
record, for every chain and iteration, the candidate window [row0, row1) and the fragment hull of
V = Z[:, k] * cnt; the chains of one UTR are then replayed step by step (bulk-synchronous, like the GPU) and
the multiply-accumulates are counted for
  alg            rows x N                         (SURVEY 8d algorithmic figure, = 1.0)
  alg_hull       rows x the chain's own hull      (lower bound with exact per-chain skipping)
  cur            what em_scan_kernel executes: 256-row blocks, sub-batches of <= 32 chains in chain order,
                 8-chain MMA columns, union hull per sub-batch
  warp32 / mi8   list sorted by KEY, warp- (32 rows) or tile-level (8 rows) row skipping per 8-chain group
  *_ghull        ... plus the fragment range of every 8-chain group instead of the sub-batch's
usage: python tests/tools/sim_scan_padding.py FIRST_UTR LAST_UTR {row|h0|hc|h1}
Result on cfg-2 UTRs 0..7 (500 reads): alg_hull 0.64, cur 1.32, warp32_ghull(hc) 1.07.  The kernel built on
that (per-group ranges, in-loop conditions) was 11 % slower than `cur` on the B200: see DESIGN.md section 5."""
import collections
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import scape_oracle as so
from scape_b200 import synth
rec=[]
orig=so._grid_argmax
cur={}
def patched(m, ch, z, k):
    K=ch.K
    lo = 0 if k==0 else int(ch.a_idx[k-1]); hi = len(m.theta)-1 if k==K-1 else int(ch.a_idx[k+1])
    v = z[:,k]*m.cnt
    nz=np.nonzero(v)[0]
    h0,h1 = (int(nz[0]), int(nz[-1])+1) if len(nz) else (0,0)
    rec.append((cur['chain'], 0, lo*13, (hi+1)*13, h0, h1, len(nz), m.n, len(m.theta)))
    return orig(m, ch, z, k)
so._grid_argmax=patched
orig_run=so.run_chain
cid=[0]
def run_chain(m,ch,rng,weights_only=False):
    cur['chain']=cid[0]; cid[0]+=1
    # wrap iteration counter by counting calls per chain
    return orig_run(m,ch,rng,weights_only)
so.run_chain=run_chain
# iteration index: count per chain
tot=collections.Counter()
KEY={'row':lambda r:(r[2],r[3]),'h0':lambda r:(r[4],r[5]),'hc':lambda r:(r[4]+r[5]),'h1':lambda r:(r[5],r[4])}[sys.argv[3]]
def simulate(recs):
    # group by step(it): recs in order per chain; assign it by order
    per=collections.defaultdict(list); cnt=collections.Counter()
    for r in recs:
        c=r[0]; it=cnt[c]; cnt[c]+=1
        per[it].append(r)
    N=recs[0][7]; T=recs[0][8]; R=T*13
    nblk=(R+255)//256
    out=collections.Counter()
    for it,rs in per.items():
        out['alg']+=sum((r[3]-r[2])*N for r in rs)            # rows*N  (MAC units)
        out['alg_hull']+=sum((r[3]-r[2])*(r[5]-r[4]) for r in rs)
        out['alg_nz']+=sum((r[3]-r[2])*r[6] for r in rs)
        for b in range(nblk):
            lo,hi=b*256,min(b*256+256,R)
            lst=[r for r in rs if r[2]<hi and r[3]>lo]     # chain order
            for first in range(0,len(lst),32):
                sub=lst[first:first+32]
                ng=(len(sub)+7)//8
                h0=min(r[4] for r in sub)&~7; h1=max(r[5] for r in sub)
                if h1<=h0: h0=h1=0
                out['cur']+=256*8*ng*((h1-h0+3)//4*4)
            # proposed: sort by row0, groups of 8, warp-level skip (32 rows), hull per sub-batch
            lst2=sorted(lst,key=KEY)
            for first in range(0,len(lst2),32):
                sub=lst2[first:first+32]
                h0=min(r[4] for r in sub)&~7; h1=max(r[5] for r in sub)
                if h1<=h0: h0=h1=0
                hl=(h1-h0+3)//4*4
                for g in range(0,len(sub),8):
                    grp=sub[g:g+8]
                    g0=min(r[2] for r in grp); g1=max(r[3] for r in grp)
                    gh0=min(r[4] for r in grp)&~3; gh1=max(r[5] for r in grp)
                    ghl=max(0,(gh1-gh0+3)//4*4)
                    for w in range(8):
                        r0=lo+32*w; r1=min(r0+32,hi)
                        if r0<g1 and r1>g0 and r0<hi:
                            out['warp32']+=32*8*hl
                            out['warp32_ghull']+=32*8*ghl
                        for mi in range(4):
                            q0=r0+8*mi; q1=q0+8
                            if q0<g1 and q1>g0 and q0<hi:
                                out['mi8']+=8*8*hl
                                out['mi8_ghull']+=8*8*ghl
    return out
for ui in range(int(sys.argv[1]), int(sys.argv[2])):
    u=synth.make_utr(ui,500)
    rec.clear(); cid[0]=0
    res=so.fit_utr(u.x,u.l,u.r,u.pa,np.random.RandomState(1))
    # only full-EM chains recorded (weights-only don't call grid)
    o=simulate(list(rec))
    tot.update(o)
    print(ui, res.n_frag, res.K, {k:round(v/o['alg'],2) for k,v in o.items()}, flush=True)
print("TOTAL", {k:round(v/tot['alg'],3) for k,v in tot.items()})
