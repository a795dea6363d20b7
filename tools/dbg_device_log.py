import numpy as np, ctypes as C, mpmath as mp, sys
sys.path.insert(0,'/root/repo')
import baynorm_b200 as bn
ctx=bn.Context(0)
mp.mp.dps=50
rng=np.random.default_rng(0)
def run(which,x):
    x=np.ascontiguousarray(x,np.float64); o=np.zeros_like(x)
    ctx.check(ctx.lib.bn_debug_log(ctx.h,which,C.c_void_p(x.ctypes.data),x.size,C.c_void_p(o.ctypes.data))); return o
x=np.concatenate([np.exp(rng.uniform(np.log(1e-300),np.log(1e300),4000)),rng.uniform(0.5,2.0,4000),1+rng.uniform(-1,1,2000)*1e-6,[1.0,0.6875,1.375,2.0,0.5,1e-3,64.0]])
got=run(0,x); ex=np.array([float(mp.log(mp.mpf(float(v)))) for v in x]); err=np.abs(got-ex)
k=np.argmax(err/(2.3e-16*np.abs(ex)+2e-18)); print('log worst', x[k], got[k], ex[k], err[k], err[k]/(2.3e-16*abs(ex[k])+2e-18))
y=np.concatenate([np.exp(rng.uniform(np.log(1e-30),np.log(1e6),6000)),rng.uniform(0,3,3000),[0.0,1.0,1e-17]])
got=run(1,y); ex=np.array([float(mp.log1p(mp.mpf(float(v)))) for v in y]); err=np.abs(got-ex)
k=np.argmax(err/(2.3e-16*np.abs(ex)+2e-18)); print('log1p worst', y[k], got[k], ex[k], err[k], err[k]/(2.3e-16*abs(ex[k])+2e-18))
ys=np.concatenate([np.exp(rng.uniform(np.log(1e-30),np.log(2.0**-6),4000)),[0.0,2.0**-6]])
got=run(2,ys); ex=np.array([float(mp.log1p(mp.mpf(float(v)))) for v in ys]); err=np.abs(got-ex)
k=np.argmax(err/(2.3e-16*np.abs(ex)+1e-300)); print('small worst', ys[k], got[k], ex[k], err[k]/(abs(ex[k])+1e-300))
z=np.concatenate([rng.uniform(64,200,2000),np.exp(rng.uniform(np.log(64),np.log(1e7),2000))])
got=run(3,z); ex=np.array([float(mp.loggamma(mp.mpf(float(v)))) for v in z]); err=np.abs(got-ex)
k=np.argmax(err/np.abs(ex)); print('stirling worst', z[k], got[k], ex[k], err[k]/abs(ex[k]))
