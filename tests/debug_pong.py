import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import pufferlib_b200.vector as pvec
from pufferlib_b200.environments import ocean
from oracle.envs import OracleVec
n=3
vec = pvec.make(ocean.env_creator('pong'), env_kwargs=dict(max_score=1,max_ticks=120), num_envs=n, backend=pvec.B200)
ora = OracleVec('pong', n, iparam=[1,120])
tape = np.random.default_rng(0).integers(0,6,size=(60,n))
vec.async_reset(5); ora.async_reset(5)
for t in range(40):
    o,r,term,_,_,_,_ = vec.recv(); oo,orr,ot,_,_,_,_ = ora.recv()
    o=o.cpu().numpy(); r=r.cpu().numpy(); term=term.cpu().numpy()
    if not np.array_equal(o,oo):
        d=np.argwhere(o!=oo)
        print('step',t,'obs mismatch count',len(d),'first',d[:8].tolist())
        e,s_,y,x=d[0]
        print(' dev vals',[int(o[tuple(k)]) for k in d[:8]],'ora vals',[int(oo[tuple(k)]) for k in d[:8]])
        for e in range(n):
            for s_ in range(4):
                print('  env',e,'slot',s_,'dev nz',np.argwhere(o[e,s_]>0)[:3].tolist(), (o[e,s_]>0).sum(),'ora nz',np.argwhere(oo[e,s_]>0)[:3].tolist(), (oo[e,s_]>0).sum())
        break
    if not np.array_equal(r,orr): print('step',t,'reward mismatch',r,orr); break
    if not np.array_equal(term,ot): print('step',t,'term mismatch',term,ot); break
    vec.send(tape[t]); ora.send(tape[t])
else:
    print('pong 40 steps ok')
