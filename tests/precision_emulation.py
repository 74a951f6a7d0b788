#!/usr/bin/env python
"""CPU emulation of reduced-precision operand schemes on the reference's encoder (test infrastructure, no GPU needed).

The fp64 oracle forward is re-run with the operands of chosen layers rounded the way a tensor-core mode would see them -- bf16, fp16,
bf16 hi+lo ("split"), TF32 (truncated or rounded to 10 mantissa bits) -- accumulation stays fp64, so only OPERAND rounding is measured.
It reproduces the all-bf16 error the CUDA path shows (3.1e-2 on the fixture), and answers what the kernels need not be built to learn:
  * activation rounding (2.8e-2) outweighs weight rounding (1.5e-2): no two-product bf16 scheme reaches the 2e-2 bound;
  * net3DV_1 layers 1-2 and the 259-wide layer must keep the split products ("bf16" mixed mode, 9e-3);
  * kind::tf32 for net3DV_3 layers 2-3 gives 1.2e-3 / 1.6e-3 on x / x_global: outside the 1e-3 fp32 bound, not built;
  * fp16 operands everywhere would give 3.9e-3 (11-bit mantissa), at the price of a range-managed backward -- not built, noted.
usage: python tests/precision_emulation.py
"""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import encoder as enc
torch.set_num_threads(8)
z=np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'train_step.npz'))
sd=oracle.init_state_dict(seed=int(z['seed_sd']))
for k in list(sd):
    if 'sd0/'+k in z.files: sd[k]=torch.from_numpy(z['sd0/'+k]).clone()
B,G,N,S,K=(int(v) for v in z['cfg'])
pts=torch.from_numpy(z['points'])
clouds=pts.permute(1,0,2,3).reshape(-1,N,4).float()
xt,yt,_=oracle.group_points(clouds,S,K,float(z['r2']))

def rb(t): return t.to(torch.bfloat16).to(torch.float64)
def rh(t): return t.to(torch.float16).to(torch.float64)
def split2(t):
    hi=rb(t); return hi+rb(t-hi)
MODES={'x':lambda t:t,'b':rb,'h':rh,'s':split2}
# mode per layer: (wmode, amode)
def forward(modes):
    sd64={k:(v.clone().double() if v.dtype.is_floating_point else v.clone()) for k,v in sd.items()}
    M=xt.shape[0]
    rows=xt.double().permute(0,2,3,1).reshape(M*S*K,4)
    h=rows
    li=0
    def mlp(h,layers,li):
        for conv,bn,ci,co in layers:
            w=sd64[conv+'.weight'].reshape(co,-1)
            wm,am=modes[li]
            zz=MODES[am](h)@MODES[wm](w).t()+sd64[conv+'.bias']
            h=torch.relu(enc._batchnorm(zz,sd64,bn,True))
            li+=1
        return h,li
    h,li=mlp(h,enc.L1_LAYERS,0)
    pooled=h.reshape(M*S,K,-1).max(dim=1).values
    centre=yt.double().reshape(M,3,S).permute(0,2,1).reshape(M*S,3)
    h,li=mlp(torch.cat([centre,pooled],1),enc.L3_LAYERS,li)
    feat=h.reshape(M,S,-1).max(dim=1).values
    Bq=M//G
    fs=feat.reshape(G,Bq,-1).max(dim=0).values
    def head(f):
        zz=f@sd64['netR_FC.0.weight'].t()+sd64['netR_FC.0.bias']
        hh=torch.relu(enc._batchnorm(zz,sd64,'netR_FC.1',True))
        return hh@sd64['netR_FC.3.weight'].t()+sd64['netR_FC.3.bias']
    return head(feat),head(fs)
def rel2(a,b): return float((a-b).norm()/b.norm())
ex=[('x','x')]*6
x0,g0=forward(ex)
print('vs fixture', rel2(x0,torch.from_numpy(z['x64'])))
def show(name,modes):
    x,g=forward(modes); print(f'{name:50s} x {rel2(x,x0):.3e} xg {rel2(g,g0):.3e}')
allb=[('s','s')]+[('b','b')]*5
show('all bf16 (layer0 split)',allb)
show('weights exact, acts bf16',[('s','s')]+[('x','b')]*5)
show('weights bf16, acts exact',[('s','s')]+[('b','x')]*5)
show('weights split, acts bf16',[('s','s')]+[('s','b')]*5)
show('weights bf16, acts split',[('s','s')]+[('b','s')]*5)
show('all fp16',[('s','s')]+[('h','h')]*5)
show('L12 fp16, rest bf16',[('s','s'),('h','h'),('h','h'),('b','b'),('b','b'),('b','b')])
show('L12,L3 fp16, L4 L5 bf16',[('s','s'),('h','h'),('h','h'),('h','h'),('b','b'),('b','b')])
show('L12,L3 split, L4 L5 bf16',[('s','s'),('s','s'),('s','s'),('s','s'),('b','b'),('b','b')])
show('L12 W split+acts bf16, L3 split',[('s','s'),('s','b'),('s','b'),('s','s'),('b','b'),('b','b')])
show('L12 W bf16+acts split, L3 split',[('s','s'),('b','s'),('b','s'),('s','s'),('b','b'),('b','b')])
show('L1 only layer1 bf16',[('s','s'),('b','b'),('s','s'),('s','s'),('s','s'),('s','s')])
show('L1 only layer2 bf16',[('s','s'),('s','s'),('b','b'),('s','s'),('s','s'),('s','s')])

def tf32_trunc(t):   # kind::tf32 reads fp32 operands and drops the low 13 mantissa bits
    a=t.to(torch.float32).contiguous().view(torch.int32) & ~0x1FFF
    return a.view(torch.float32).to(torch.float64)
def tf32_rn(t):
    a=t.to(torch.float32).contiguous().view(torch.int32)
    a=(a + 0x1000) & ~0x1FFF
    return a.view(torch.float32).to(torch.float64)
MODES['t']=tf32_trunc; MODES['r']=tf32_rn
def show(name,modes):
    x,g=forward(modes); print(f'{name:58s} x {rel2(x,x0):.3e} xg {rel2(g,g0):.3e}')
sp=('s','s')
show('fp32 mode (all split)',[sp]*6)
show('L4,L5 tf32 (truncating), rest split',[sp,sp,sp,sp,('t','t'),('t','t')])
show('L4,L5 tf32 (operands pre-rounded to nearest), rest split',[sp,sp,sp,sp,('r','r'),('r','r')])
show('L5 tf32 truncating only',[sp,sp,sp,sp,sp,('t','t')])
show('all layers tf32 (rn)',[sp]+[('r','r')]*5)
