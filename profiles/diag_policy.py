import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R+'/tests'); sys.path.insert(0,R)
import torch, numpy as np
import test_gpu_policy as tg
n,T=640,20
eng,o=tg.make_engine(n,T); eng.collect(); b=eng.buf
stacks=tg.oracle_stacks(b,T,n)
logits=torch.zeros((n,7),device='cuda'); val=torch.zeros(n,device='cuda'); age=torch.zeros(n,dtype=torch.uint8,device='cuda')
for t in range(T+1):
    pa=None if t==0 else b['age'][t-1]; pd=None if t==0 else b['ep_len'][t-1]
    eng.policy.forward_rollout(b['frames'],b['dirs'],b['mission'][t+3],t+3,pa,pd,age,val,logits=logits)
    img,d,mis=stacks[t]
    with torch.no_grad(): lo,vo=o({'direction':torch.from_numpy(d),'image':torch.from_numpy(img),'mission':torch.from_numpy(mis)})
    # also torch-GPU evaluate of product network
    dl=(logits.cpu()-lo).abs(); dv=(val.cpu()-vo).abs()
    print(t,'logits maxabs %.2e rel-to-max %.2e | value maxabs %.2e rel %.2e | elemwise rel max %.2e'%(dl.max(),dl.max()/lo.abs().max(),dv.max(),dv.max()/vo.abs().max(),(dl/(lo.abs()+1e-12)).max()))
