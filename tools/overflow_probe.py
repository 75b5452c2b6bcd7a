import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mujoco_manip_b200 import PickPlaceVecEnv
dev=torch.device('cuda:0'); n=4096
env=PickPlaceVecEnv(n, device=dev, task=("obj_red","bin_red"), seed=1234)
env.reset()
gen=torch.Generator(device=dev).manual_seed(1234)
T0=env.state["tinit"][0]; p0,R0=T0[:3],T0[3:].reshape(3,3)
lo=torch.tensor([-0.3,0.30,0.30],device=dev,dtype=torch.float64); hi=torch.tensor([0.3,0.65,0.60],device=dev,dtype=torch.float64)
ov=torch.zeros(n,dtype=torch.int32,device=dev); mx=0
for t in range(120):
    w=lo+(hi-lo)*torch.rand((n,3),device=dev,dtype=torch.float64,generator=gen)
    a=torch.zeros((n,10),device=dev); a[:,:3]=((w-p0)@R0).float(); a[:,6]=1; a[:,7]=(torch.rand(n,device=dev,generator=gen)>0.5).float()
    env.step(a)
    ov |= env.state["diag"][:,2]
    mx=max(mx,int(env.state["diag"][:,0].max()))
print("envs with any overflow bit over 120 steps:", int((ov!=0).sum()), "bits:", [int(((ov>>b)&1).sum()) for b in range(3)], "max ncon", mx, "nonfinite resets", int(env.state["diag"][:,3].sum()))
