import sys; sys.path.insert(0,'.')
import torch
from balance_robot_b200 import make_vec
for wb in (False, True):
    n=65536
    env=make_vec("Env03-v2", n, seed=0, wheel_block=wb); env.reset()
    gen=torch.Generator(device="cuda").manual_seed(1234)
    acts=[torch.rand((n,2),device="cuda",generator=gen)*2-1 for _ in range(8)]
    for k in range(30): env.step(acts[k%8])
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(10): env.step(acts[k%8])
    e1.record(); torch.cuda.synchronize()
    print('wheel_block',wb,'ms/step %.2f'%(e0.elapsed_time(e1)/10), {k:v for k,v in env.stats().items() if k in ('unsupported','coupled_substeps','coupled_fallbacks')})
    env.close()
