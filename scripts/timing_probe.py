import os, sys, time, io, contextlib
os.environ["FITBURST_B200_TIMING"]="1"
sys.path.insert(0,'/root/repo')
import numpy as np, torch, bench
from fitburst_b200 import synthetic as syn
from fitburst_b200.analysis.fitter import LSFitter
from fitburst_b200.analysis.model import SpectrumModeler
class A: nfreq=16384; ntime=1024; ncomp=3; kf=8; kt=4
args=A()
freqs,times=syn.chime_axes(args.nfreq,args.ntime)
model=SpectrumModeler(freqs,times,factor_freq_upsample=8,factor_time_upsample=4,num_components=3)
def model_fn(p):
    model.update_parameters(p); return model.compute_model()
b=bench.make_burst(args,0,model_fn)
with contextlib.redirect_stdout(io.StringIO()):
    model.update_parameters(b[1]); f=LSFitter(b[3],model,b[2]); f.fix_parameter(["dm_index","scattering_index"])
for r in range(5):
    model.update_parameters(b[1]); torch.cuda.synchronize(); t0=time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        f._fit_native(f.get_fit_parameters_list())
    t1=time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        f.results=None; h,_=f.compute_hessian(f.data,f.fit_parameters)
    t2=time.perf_counter()
    print(f"fit_native {1e3*(t1-t0):.2f} ms  hessian {1e3*(t2-t1):.2f} ms", file=sys.stderr)
