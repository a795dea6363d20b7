import sys, numpy as np
sys.path.insert(0, '/root/repo')
import baynorm_b200 as bn
from oracle import oracle as O
fx = O.load_fixture()
X, beta, size, mu = (fx[k] for k in ("inputdata", "inputbeta", "size", "mu"))
ctx = bn.Context(0)
M = bn.Check_input(X, ctx); Mo = O.CSC.from_dense(X)
S = 3000
d = bn.Main_NB_Bay(M, beta, size, mu, S=S, seed=12345)
nb_size, nb_mu = O.post_sample_params(Mo, beta, size, mu)
mean_th = nb_mu + X; var_th = nb_mu + nb_mu**2/nb_size
z = (d.mean(axis=2) - mean_th)/np.sqrt(var_th/S)
vz = d.var(axis=2, ddof=1)/var_th
order = np.argsort(-np.abs(z).ravel())[:15]
for o in order:
    i, j = np.unravel_index(o, z.shape)
    print("gene %2d cell %2d x=%4.0f r=%7.3f M=%8.3f  z=%7.2f  mean got %.3f th %.3f  var ratio %.3f" % (i, j, X[i,j], nb_size[i,j], nb_mu[i,j], z[i,j], d[i,j].mean()-X[i,j], nb_mu[i,j], vz[i,j]))
light = nb_mu < 8
print("light: max|z| %.2f  heavy: max|z| %.2f  n_heavy %d" % (np.abs(z[light]).max(), np.abs(z[~light]).max(), (~light).sum()))
