"""Randomised shapes / precisions / epilogues for the experimental conv2 kernel on the CPU model of the primitives
(tests/sim).  usage: python tools/conv2_fuzz.py [seed] [cases]    (builds nothing: make -C piper_b200/csrc first)"""
import sys, ctypes as C, numpy as np, torch, time
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import test_conv2_sim as T
sim=C.CDLL(T.SIM)
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
n=int(sys.argv[2]) if len(sys.argv)>2 else 30
bad=0
for it in range(n):
    prec=int(rng.integers(0,3))
    ci=int(rng.choice([16,32,48,64,96,128,192,256])); 
    if prec==1 and ci%8: continue
    rows=int(rng.choice([16,32,48,64,96,128,192,256,384]))
    k=int(rng.choice([1,3,5,7])); dil=int(rng.choice([1,2,3,6])) if k>1 else 1
    B=int(rng.integers(1,4)); lens=tuple(int(x) for x in rng.integers(1,420,size=B))
    epi=str(rng.choice(["BIAS","RES","RELU"])); pre=int(rng.integers(0,2)); chains=int(rng.integers(1,3))
    grid=int(rng.integers(1,4))
    x,clean=T._ragged(B,ci,lens,seed=it)
    w=(rng.standard_normal((rows,ci,k))/np.sqrt(ci*k)).astype(np.float32); bias=rng.standard_normal(rows).astype(np.float32)
    r=rng.standard_normal((B,rows,x.shape[2])).astype(np.float32)
    t=time.time()
    try:
        y,_,info=T._run(sim,x,w,bias,lens,dil=dil,pre=pre,epi=epi,prec=prec,chains=chains,r=r if epi=="RES" else None,grid=grid)
    except AssertionError as e:
        print("FAIL",(prec,ci,rows,k,dil,lens,epi,pre,chains,grid),str(e)[:200]); bad+=1; continue
    worst=0
    for b,L in enumerate(lens):
        ref=T._ref_conv(clean[b],w,bias,dil,(k-1)//2*dil,pre,0.1)
        if epi=="RELU": ref=torch.relu(ref)
        if epi=="RES": ref=ref+torch.from_numpy(r[b,:,:L])
        e=float((torch.from_numpy(y[b,:,:L])-ref).abs().max())/max(1.0,float(ref.abs().max()))
        worst=max(worst,e)
        if not np.all(y[b,:,L:]==7e7): print("WROTE OUTSIDE",(prec,ci,rows,k,dil,lens)); bad+=1
    tol=3e-4 if prec==0 else 2e-5
    flag="" if worst<=tol else "  <-- BAD"
    if flag: bad+=1
    print(f"prec={prec} ci={ci:3d} rows={rows:3d} k={k} d={dil} lens={lens} {epi:4s} pre={pre} chains={chains} grid={grid} plan(n_tile,n_tiles,mt,kc,slots)={info[:5]} err={worst:.1e} {time.time()-t:.1f}s{flag}",flush=True)
print("bad:",bad)
