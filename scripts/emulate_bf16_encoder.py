"""CPU emulation of the bf16 mini-PointNet pipeline (same rounding points as act_b200/layers.py PointNetEncoderFn)\nagainst the reference golden gradients: shows the deep-gradient error is a property of bf16 compute."""
import numpy as np, torch, sys
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT)
from oracle import ref_model
torch.set_num_threads(8)
g=np.load(os.path.join(ROOT,'tests','golden','')+'encoder.npz'); grp=np.load(os.path.join(ROOT,'tests','golden','')+'group.npz')
nb=torch.from_numpy(grp['shapenet/neighborhood'][:2])
enc=ref_model.fill_params(ref_model.Encoder(384),seed=2).train()
sd=enc.state_dict()
r=lambda t: t.bfloat16().float()
def run(rnd):
    q = r if rnd else (lambda t:t)
    B,G,k,_=nb.shape; M=B*G*k
    p=nb.reshape(M,3)
    W1=sd['first_conv.0.weight'].view(128,3); b1=sd['first_conv.0.bias']; g1=sd['first_conv.1.weight']; be1=sd['first_conv.1.bias']
    W2=q(sd['first_conv.3.weight'].view(256,128)); b2=sd['first_conv.3.bias']
    W3=q(sd['second_conv.0.weight'].view(512,512)); b3=sd['second_conv.0.bias']; g2=sd['second_conv.1.weight']; be2=sd['second_conv.1.bias']
    W4=q(sd['second_conv.3.weight'].view(384,512)); b4=sd['second_conv.3.bias']
    h1=p@W1.t()+b1; m1=h1.mean(0); v1=h1.var(0,unbiased=False); rs1=torch.rsqrt(v1+1e-5)
    xh1=(h1-m1)*rs1
    a1=q(torch.relu(xh1*g1+be1))
    f2f=a1@W2.t()+b2; f2=q(f2f)
    gmaxf,arg2=f2f.view(B*G,k,256).max(1); gmax=q(gmaxf)
    gpart=gmax@W3[:,:256].t()+b3
    h3=q(f2@W3[:,256:].t()+gpart.repeat_interleave(k,0))
    m2=h3.mean(0); v2=h3.var(0,unbiased=False); rs2=torch.rsqrt(v2+1e-5)
    a3=q(torch.relu((h3-m2)*rs2*g2+be2))
    f4=a3@W4.t()+b4
    tok,arg4=f4.view(B*G,k,384).max(1)
    d=torch.from_numpy(g['wout']).view(B*G,384)
    dF4=torch.zeros(B*G,k,384).scatter_(1,arg4[:,None],q(d)[:,None]).view(M,384)
    dZ3=q((dF4@W4)*(a3>0))
    xh3=(h3-m2)*rs2
    s1=dZ3.sum(0); s2=(dZ3*xh3).sum(0)
    dH3=q(g2*rs2*(dZ3-s1/M-xh3*s2/M))
    dGp=dH3.view(B*G,k,512).sum(1); dGpb=q(dGp)
    dgmax=dGpb@W3[:,:256]
    dF2=q(dH3@W3[:,256:])
    dF2=q(dF2+torch.zeros(B*G,k,256).scatter_(1,arg2[:,None],dgmax[:,None]).view(M,256))
    dZ1=q((dF2@W2)*(a1>0))
    dg1=(dZ1*xh1).sum(0); dbe1=dZ1.sum(0)
    dh1=g1*rs1*(dZ1-dbe1/M-xh1*dg1/M)
    dW1=dh1.t()@p
    return tok,dg1,dW1
rel=lambda a,b:((a-b).norm()/b.norm()).item()
for rnd in (False,True):
    tok,dg1,dW1=run(rnd)
    print(rnd, rel(tok.view(2,64,384),torch.from_numpy(g['out'])), rel(dg1,torch.from_numpy(g['grad/first_conv.1.weight'])), rel(dW1,torch.from_numpy(g['grad/first_conv.0.weight']).view(128,3)))
