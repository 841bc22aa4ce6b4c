import sys, numpy as np
sys.path.insert(0,'.')
from trep_b200 import lib, systems
def t_step(name,B,ns):
    d=systems.named_desc(name); s=lib.System(d)
    rng=np.random.default_rng(0)
    q=rng.uniform(-3,3,(B,d.nq)); up=lambda a: lib.DeviceBuffer(0,a.shape,a.dtype).upload(a)
    dq=up(q); dp=lib.DeviceBuffer(0,(B,d.nd)); s.calc_p2_raw(True,B,0.01,dq,dq,dp)
    q2=lib.DeviceBuffer(0,(B,d.nq)); p2=lib.DeviceBuffer(0,(B,d.nd)); it=lib.DeviceBuffer(0,(B,),np.int32); st=lib.DeviceBuffer(0,(B,),np.int32)
    ms=[]
    for r in range(4):
        s.step_raw(True,B,ns,0.01,0.01,dq,dp,None,None,None,None,q2,p2,None,it,st); lib.synchronize(0); ms.append(s.last_kernel_ms())
    t=np.mean(ms[1:]); print("%-18s %-18s steps/s %.4g  regs %s"%(name,s.kernel_name,B*ns/t*1e3,s.kernel_info(0)))
t_step("damped_pendulum",1<<20,1000)
t_step("pendulum1",1<<20,1000)
t_step("dual_pendulums",1<<22,100)
t_step("pendulum5",1<<18,200)
t_step("fourbar",1<<18,100)

