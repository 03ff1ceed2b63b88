"""Randomised runs of the experimental conv2 kernel on the CPU model of the primitives, fused epilogues: WaveNet gate, WN
residual / skip, the three MRF modes, ConvTranspose pixel-shuffle (strides 8 / 4 / 2), coupling subtract.
usage: python tools/conv2_fuzz_epilogues.py <seed> <cases>      (make -C piper_b200/csrc first)"""
import sys, os, ctypes as C, numpy as np, torch, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import test_conv2_sim as T
import torch.nn.functional as F
sim=C.CDLL(T.SIM)
rng=np.random.default_rng(int(sys.argv[1])); n=int(sys.argv[2]); bad=0
for it in range(n):
    prec=int(rng.integers(0,3)); epi=str(rng.choice(["GATE","WN","MRF","UPSAMPLE","SUBFROM"]))
    B=int(rng.integers(1,4)); lens=tuple(int(x) for x in rng.integers(1,300,size=B)); grid=int(rng.integers(1,4)); chains=int(rng.integers(1,3))
    tol=3e-4 if prec==0 else 2e-5
    try:
        if epi=="UPSAMPLE":
            up,ku=[(8,16),(4,8),(2,4)][int(rng.integers(0,3))]; ci=int(rng.choice([32,64,128])); co=int(rng.choice([16,32,64]))
            x,clean=T._ragged(B,ci,lens,seed=it); Wt=(rng.standard_normal((ci,co,ku))/np.sqrt(ci*2)).astype(np.float32); bc=rng.standard_normal(co).astype(np.float32)
            m=ku//up; w=np.zeros((co*up,ci,m),np.float32)
            for j in range(m):
                for phi in range(up): w[phi::up,:,j]=Wt[:,:,phi+(m-1-j)*up].T
            y,_,info=T._run(sim,x,w,np.repeat(bc,up),lens,pad=m-1,pre=1,epi=epi,up=up,up_pad=up//2,q_extra=m-1,y_channels=co,prec=prec,chains=chains,grid=grid)
            worst=0
            for b,L in enumerate(lens):
                ref=F.conv_transpose1d(F.leaky_relu(clean[b],0.1)[None].double(),torch.from_numpy(Wt).double(),torch.from_numpy(bc).double(),stride=up,padding=up//2)[0].float()
                worst=max(worst,float((torch.from_numpy(y[b,:,:L*up])-ref).abs().max())/max(1,float(ref.abs().max())))
                assert np.all(y[b,:,L*up:]==7e7)
            desc=(ci,co,up)
        else:
            ci=int(rng.choice([16,32,64,96,192])); 
            if prec==1 and ci%8: continue
            k=int(rng.choice([1,3,5])); dil=int(rng.choice([1,2,3])) if k>1 else 1
            x,clean=T._ragged(B,ci,lens,seed=it)
            if epi=="GATE":
                H=int(rng.choice([16,32,64,96])); rows=2*H
            elif epi=="WN":
                H=int(rng.choice([32,64,96])); rows=2*H if rng.integers(0,2) else H
            else:
                rows=int(rng.choice([16,32,64,128])); H=rows
            w=(rng.standard_normal((rows,ci,k))/np.sqrt(ci*k)).astype(np.float32); bias=rng.standard_normal(rows).astype(np.float32)*0.3
            cs=x.shape[2]; worst=0
            conv=lambda b:T._ref_conv(clean[b],w,bias,dil,(k-1)//2*dil,0,0.1)
            if epi=="GATE":
                y,_,info=T._run(sim,x,w,bias,lens,dil=dil,epi=epi,prec=prec,chains=chains,grid=grid,y_channels=H)
                for b,L in enumerate(lens):
                    p=conv(b); ref=torch.tanh(p[0::2])*torch.sigmoid(p[1::2]); worst=max(worst,float((torch.from_numpy(y[b,:,:L])-ref).abs().max()))
                tol=max(tol,2e-5)
            elif epi=="SUBFROM":
                r=rng.standard_normal((B,rows,cs)).astype(np.float32)
                y,_,info=T._run(sim,x,w,bias,lens,dil=dil,epi=epi,prec=prec,chains=chains,grid=grid,r=r)
                for b,L in enumerate(lens):
                    ref=torch.from_numpy(r[b,:,:L])-conv(b); worst=max(worst,float((torch.from_numpy(y[b,:,:L])-ref).abs().max())/max(1,float(ref.abs().max())))
            elif epi=="WN":
                split=H if rows==2*H else 0; first=int(rng.integers(0,2))
                r=rng.standard_normal((B,max(split,1),cs)).astype(np.float32); sk=rng.standard_normal((B,rows-split,cs)).astype(np.float32)
                y,y2,info=T._run(sim,x,w,bias,lens,dil=dil,epi=epi,prec=prec,chains=chains,grid=grid,r=r,y2_init=sk,split=split,first=first,y_channels=max(split,1))
                for b,L in enumerate(lens):
                    v=conv(b)
                    if split: worst=max(worst,float((torch.from_numpy(y[b,:,:L])-(torch.from_numpy(r[b,:,:L])+v[:split])).abs().max())/max(1,float(v.abs().max())))
                    want=v[split:] if first else torch.from_numpy(sk[b,:,:L])+v[split:]
                    worst=max(worst,float((torch.from_numpy(y2[b,:,:L])-want).abs().max())/max(1,float(want.abs().max())))
            else: # MRF
                if rows!=ci: rows=ci; w=(rng.standard_normal((rows,ci,k))/np.sqrt(ci*k)).astype(np.float32); bias=rng.standard_normal(rows).astype(np.float32)*0.3; conv=lambda b:T._ref_conv(clean[b],w,bias,dil,(k-1)//2*dil,0,0.1)
                if rows%16: continue
                mode=int(rng.integers(0,3)); acc0=rng.standard_normal((B,rows,cs)).astype(np.float32)
                _,y2,info=T._run(sim,x,w,bias,lens,dil=dil,epi=epi,prec=prec,chains=chains,grid=grid,r=x,y2_init=acc0,mrf=mode,mrf_n=3)
                for b,L in enumerate(lens):
                    v=conv(b)+clean[b]; a0=torch.from_numpy(acc0[b,:,:L]); want=v if mode==0 else (a0+v if mode==1 else (a0+v)/3)
                    worst=max(worst,float((torch.from_numpy(y2[b,:,:L])-want).abs().max())/max(1,float(want.abs().max())))
            desc=(ci,rows,k,dil)
    except AssertionError as e:
        print("FAIL",epi,prec,lens,str(e)[:150]); bad+=1; continue
    flag="" if worst<=tol else "  <-- BAD"
    bad+= bool(flag)
    print(f"{epi:8s} prec={prec} {desc} lens={lens} chains={chains} grid={grid} plan={info[:5]} err={worst:.1e}{flag}",flush=True)
print("bad:",bad)
