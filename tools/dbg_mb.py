import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
import howl_b200
from howl_b200 import mobilenet as mb
from oracle import howl_oracle as O
from test_gpu_mobilenet import _random_state, _flat_of, _bn_of
DEV=torch.device('cuda:0')
ctx=howl_b200.Context('cuda:0', n_mels=40)
L=12
for B in (4, 16, 64, 256):
    sd=_random_state(L, seed=B)
    pcm,_=O.synthetic_batch(B,16000,L,seed=B+1)
    fb=O.mel_filterbank(40); zm=(-2.0166,3.9955)
    feats=ctx.frontend(pcm.to(DEV), fb.to(DEV), "mels", zmuv=zm)
    x=O.hot_path_features(pcm,fb,torch.tensor([zm[0]]),torch.tensor([zm[0]**2+zm[1]**2]))
    for train in (False, True):
        bn=_bn_of(sd,L).to(DEV); nbt=torch.zeros(mb.bn_layers(ctx),dtype=torch.int64,device=DEV)
        ws=torch.empty(mb.workspace_bytes(ctx,B,feats.shape[2],L),dtype=torch.uint8,device=DEV)
        logits=mb.forward(ctx,feats,_flat_of(sd,L).to(DEV),bn,nbt,train,ws).cpu()
        with torch.no_grad(): want=O.mobilenet_forward(x,sd,train=train)
        # oracle with bf16-rounded activations emulation is not available; report rel error
        print(B, train, ((logits-want).norm()/want.norm()).item(), want.abs().mean().item())
