"""Debug aid kept from the RAFT fused lookup+convc1 investigation ("things tried"): runs the kernel in the
bf16x3 / bf16 / fp16 engines against relu(conv1x1(oracle lookup)) and, for the single-MMA path, fits which K-steps and
which pyramid levels the result contains (a dropped K-step shows up as a level coefficient near 0).
    python tools/experiments/dbg_raft_fused.py      (GPU)
"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import anystereo_b200 as A, cases
from oracle import hotpath_oracle as O
B,H,W,Lv=2,5,23,4
c=cases.raft_corr_case(seed=41+H,B=B,D=16,H=H,W=W,L=Lv)
rng=np.random.RandomState(W+Lv)
disp=torch.from_numpy(rng.uniform(-8,W+8,(B,1,H,W)).astype('float32'))
coords=torch.arange(W).float().reshape(1,1,W,1).repeat(B,H,1,1)
w=torch.from_numpy(rng.standard_normal((64,Lv*9,1,1)).astype('float32'))*0.3
b=torch.from_numpy(rng.standard_normal(64).astype('float32'))*0.1
feat=O.corrblock1d_lookup(O.corr_pyramid(O.all_pairs_corr(c['f1'],c['f2']),Lv),disp,coords,4,exact=True)
ref=torch.relu(torch.nn.functional.conv2d(feat.double(),w.double(),b.double())).float()
A.set_corr_mode('fp32')
for engine in ('bf16x3','bf16','fp16'):
    A.set_update_engine(engine)
    blk=A.CorrBlock1D(c['f1'].cuda(),c['f2'].cuda(),num_levels=Lv,radius=4)
    split=engine=='bf16x3'
    d=blk.deferred(disp.cuda(),coords.cuda())
    w_hi,w_lo=type(d).pack_convc1_weight(w.cuda(),split)
    widen=(lambda t:t.view(torch.float16).float()) if engine=='fp16' else (lambda t:t.float())
    # check packed weights
    wp=widen(w_hi).cpu()
    cidx=torch.arange(36); k=(cidx//9)*10+cidx%9
    print(engine,'weight pack err',float((wp[:,k]-w.reshape(64,36)).abs().max()), 'pad max', float(wp[:, [9,19,29,39]+list(range(40,64))].abs().max()))
    for rep in range(2):
        out_hi=torch.full((B,H,W,64),float('nan'),device='cuda',dtype=torch.bfloat16)
        out_lo=torch.full_like(out_hi,float('nan')) if split else None
        d.convc1_planes(w_hi,w_lo,b.cuda(),out_hi,out_lo)
        torch.cuda.synchronize()
        got=(widen(out_hi)+(widen(out_lo) if split else 0)).permute(0,3,1,2).cpu()
        err=(got-ref).abs()
        print(engine,rep,'max err',float(err.max()),'ref max',float(ref.abs().max()),'n bad',int((err>0.05).sum()),'of',err.numel())
        bad=(err>0.05).nonzero()
        if len(bad):
            print('  bad channels',sorted(set(bad[:,1].tolist()))[:20],' bad pixels(flat)',sorted(set((bad[:,0]*H*W+bad[:,2]*W+bad[:,3]).tolist()))[:20])
A.set_update_engine('fp32')

# which K-steps (16 K values each) does the nsplit=1 result contain?
import itertools
A.set_update_engine('bf16')
blk=A.CorrBlock1D(c['f1'].cuda(),c['f2'].cuda(),num_levels=Lv,radius=4)
d=blk.deferred(disp.cuda(),coords.cuda())
w_hi,_=type(d).pack_convc1_weight(w.cuda(),False)
out_hi=torch.zeros((B,H,W,64),device='cuda',dtype=torch.bfloat16)
d.convc1_planes(w_hi,None,torch.zeros(64,device='cuda'),out_hi,None)     # zero bias; relu still applied
torch.cuda.synchronize()
got=out_hi.float().cpu().reshape(-1,64)
F=torch.zeros(B*H*W,64)
f36=feat.permute(0,2,3,1).reshape(-1,36)
cidx=torch.arange(36); kk=(cidx//9)*10+cidx%9
F[:,kk]=f36
Wp=torch.zeros(64,64); Wp[:,kk]=w.reshape(64,36)
for r in range(1,5):
    for sub in itertools.combinations(range(4),r):
        cols=[i for s_ in sub for i in range(16*s_,16*s_+16)]
        pred=torch.relu(F[:,cols]@Wp[:,cols].t())
        e=float((pred-got).abs().max())
        if e<0.3: print('k-steps',sub,'max err',e)
A.set_update_engine('fp32')
refz=torch.relu(F@Wp.t())
print('zero-bias check: max err vs full-K pred', float((refz-got).abs().max()))
idx=(refz-got).abs().flatten().topk(6).indices
for i in idx.tolist():
    p,ch=divmod(i,64)
    print('pixel',p,'ch',ch,'got',float(got[p,ch]),'ref',float(refz[p,ch]))
# per-level partial sums: which level's contribution is missing/wrong?
for lv_ in range(4):
    cols=list(range(lv_*10,lv_*10+10))
    part=F[:,cols]@Wp[:,cols].t()
    print('level',lv_,'|partial| mean',float(part.abs().mean()))
# least squares: got_pre ~ sum_l a_l * partial_l on entries where got>0
mask=(got>0)&(refz>0)
Pl=torch.stack([(F[:,list(range(l*10,l*10+10))]@Wp[:,list(range(l*10,l*10+10))].t())[mask] for l in range(4)],1)
sol=torch.linalg.lstsq(Pl,got[mask].unsqueeze(1)).solution.flatten()
print('lstsq level coefficients',sol.tolist())
