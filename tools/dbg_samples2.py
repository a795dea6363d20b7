import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import baynorm_b200 as bn
from oracle import oracle as O
from util import synth_counts
X, mu, size, beta = synth_counts(300, 500, seed=11)
X, beta, size, mu = X[:60, :80], beta[:80], size[:60], mu[:60] * 40
ctx = bn.Context(0)
M = bn.Check_input(X, ctx); Mo = O.CSC.from_dense(X)
S = 800
d = bn.Main_NB_Bay(M, beta, size, mu, S=S, seed=12345)
nb_size, nb_mu = O.post_sample_params(Mo, beta, size, mu)
mean_th = nb_mu + X; var_th = nb_mu + nb_mu**2/nb_size
kap4 = var_th*(1+6*var_th/nb_size)
vz = (d.var(axis=2, ddof=1) - var_th)/np.sqrt((kap4 + 2*var_th**2)/S + 1e-300)
order = np.argsort(-np.abs(vz).ravel())[:8]
for o in order:
    i, j = np.unravel_index(o, vz.shape)
    k = d[i,j]-X[i,j]
    print("gene %2d cell %2d x=%4.0f r=%8.4f M=%9.3f  vz=%6.2f var got %.3f th %.3f  max draw %.0f  sorted top %s" % (i, j, X[i,j], nb_size[i,j], nb_mu[i,j], vz[i,j], d[i,j].var(ddof=1), var_th[i,j], k.max(), np.sort(k)[-4:]))
