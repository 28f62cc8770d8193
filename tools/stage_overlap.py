import os,sys,ctypes
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import bench
from peanut_b200.pipeline import PerceptionPipeline
from peanut_b200 import _lib
wa,wc=bench.synth_weights(24)
pipe=PerceptionPipeline(wa,wc,num_envs=1,device='cuda:0',precision='tf32',map_shape=(24,240,240))
host=bench.synth_inputs(1,(24,240,240),0)
d={k:v.cuda() for k,v in host.items()}
maps,poses=d['maps'].clone(),d['poses'].clone()
flush=torch.empty(256<<20,dtype=torch.uint8,device='cuda')
a=pipe.args
def ab():
    global maps
    pipe.seg.forward_device(d['rgb'],None,a.sem_pred_prob_thr,a.sem_pred_prob_thr,a.goal_thr,out=pipe.sem)
    st=torch.cuda.current_stream().cuda_stream
    _lib.check(pipe.seg.ctx.lib.pn_make_obs(pipe.seg.ctx.handle,d['depth'].data_ptr(),d['rgb'].data_ptr(),pipe.sem.data_ptr(),1,a.env_frame_height,a.env_frame_width,a.frame_height,a.frame_width,a.num_sem_categories,a.min_depth,a.max_depth,pipe.obs.data_ptr(),ctypes.c_void_p(st)))
    fp,maps,_=pipe.mapper.forward_batch(pipe.obs,d['delta'],maps,poses)
def full():
    global maps
    _,_,maps,_,_=pipe.step_device(d['rgb'],d['depth'],d['delta'],maps,poses,d['pmap'])
def conly():
    pipe.pred.forward_device(d['pmap'],apply_sigmoid=True,out=pipe.pred_out)
for name,fn in (('A+B',ab),('full',full),('C',conly)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(20):
        flush.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(name, round(sum(ts)/len(ts),3),'ms')
