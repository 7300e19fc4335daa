import sys, time, io, contextlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from fitburst_b200 import synthetic as syn
from fitburst_b200.analysis.fitter import LSFitter
from fitburst_b200.analysis.model import SpectrumModeler
class A: nfreq=16384; ntime=1024; ncomp=3; kf=8; kt=4
args=A()
freqs,times=syn.chime_axes(args.nfreq,args.ntime)
model=SpectrumModeler(freqs,times,factor_freq_upsample=8,factor_time_upsample=4,num_components=3)
def model_fn(p):
    model.update_parameters(p); return model.compute_model()
bursts=[bench.make_burst(args,k,model_fn) for k in range(2)]
pinned=[torch.from_numpy(b[3]).pin_memory() for b in bursts]; host=[p.numpy() for p in pinned]
keep={}
def step(k, tag):
    truth,start,good,_=bursts[k]
    torch.cuda.synchronize(); t0=time.perf_counter()
    model.update_parameters(start)
    with contextlib.redirect_stdout(io.StringIO()):
        f=LSFitter(host[k],model,good); torch.cuda.synchronize(); t1=time.perf_counter()
        f.fix_parameter(["dm_index","scattering_index"]); f.fit(); torch.cuda.synchronize(); t2=time.perf_counter()
    keep["f"]=f
    print(f"{tag} burst{k}: ctor {1e3*(t1-t0):.1f} ms  fit {1e3*(t2-t1):.1f} ms nfev {f.results.nfev}")
for rep in range(4): step(0,"same")
for rep in range(6): step(rep%2,"alt ")
s=bench.ClockSampler(0); s.start()
for rep in range(6): step(rep%2,"alt+sampler")
s.stop_flag.set()
print("---- finer: alloc vs copy")
import gc
for rep in range(12):
    k=rep%2
    torch.cuda.synchronize(); t0=time.perf_counter()
    buf=torch.empty((16384,1024),dtype=torch.float64,device='cuda'); torch.cuda.synchronize(); t1=time.perf_counter()
    buf.copy_(torch.from_numpy(host[k]), non_blocking=False); torch.cuda.synchronize(); t2=time.perf_counter()
    keep["b"]=buf
    print(f"alloc {1e3*(t1-t0):.2f} ms copy {1e3*(t2-t1):.2f} ms  reserved {torch.cuda.memory_reserved()/2**20:.0f} MiB")
