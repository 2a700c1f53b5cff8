"""DEBUG TOOLING (not a test, not shipped): random synthetic workloads (coverage, error rate, low-quality bases, true
pairs, STR density, variant density, read length, first k) through the one-thread build of the device sources
(tests/hostsim) and through the compiled reference (oracle/_ref/ref_windows); every Variant_t tuple must agree and no
window may come back unassembled.  usage: python tools/sim_fuzz.py [seed] [cases]
(round 2: 240 cases, no disagreement; 4 of them ran into a capacity and were reported as not assembled: three the
documented device limit -- more than 12 288 k-mers in one (window, k): 1 % errors at k >= 63, or 240x depth with 0.6 %
errors --, one the simulation's own 64-record output slab, which the device replaces by its 1024-record slabs)"""
import os, sys, random, subprocess, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'oracle'))
import run_ref
from lancet_b200.synth import make_batch
SIM=os.path.join(ROOT,'tests','hostsim','_build','hostsim')
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
N=int(sys.argv[2]) if len(sys.argv)>2 else 50
bad=0
for it in range(N):
    kw=dict(seed=rng.randrange(1,10**6), region_len=rng.choice([800,1200,2000,3000]))
    if rng.random()<0.5: kw['err']=rng.choice([0.001,0.002,0.004,0.006])
    if rng.random()<0.3: kw['low_qual_frac']=rng.choice([0.01,0.04,0.08])
    if rng.random()<0.35:
        kw['paired']=True; kw['insert_mean']=rng.choice([130,160,220,300]); kw['insert_sd']=rng.choice([10,30,50])
    if rng.random()<0.35: kw['str_every']=rng.choice([100,150,200,400])
    if rng.random()<0.5: kw['var_every']=rng.choice([50,100,200,400,1000])
    if rng.random()<0.4:
        c=rng.choice([10,20,40,80,120]); kw['cov_t']=c; kw['cov_n']=rng.choice([c, max(5,c//2)])
    if rng.random()<0.25: kw['read_len']=rng.choice([75,125,150,250])
    mink=rng.choice([11,11,11,13,17,25,31,33,47,63,65])
    try:
        b=make_batch(**kw)
    except Exception as e:
        print('gen fail', kw, e); continue
    with tempfile.TemporaryDirectory() as td:
        p=os.path.join(td,'b.lb2b'); b.save(p)
        want,_=run_ref.run(path=p, threads=8, min_k=mink)
        out=os.path.join(td,'o.tsv')
        r=subprocess.run([SIM,p,'--out',out,'--min-k',str(mink)],capture_output=True,text=True)
        got=run_ref.parse_tsv(open(out).read()) if r.returncode==0 else None
    unassembled=[l for l in r.stderr.splitlines() if 'status 3' in l or 'status 4' in l]
    ok = got==want
    if not ok or unassembled:
        bad+=1
        print('BAD' if not ok else 'UNASM', mink, json.dumps(kw), len(want), None if got is None else len(got), unassembled[:2], flush=True)
    elif it%10==0:
        print('ok', it, mink, json.dumps(kw), len(want), flush=True)
print('done', N, 'bad', bad)
