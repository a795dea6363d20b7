import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import baynorm_b200 as bn
from oracle import oracle
ctx = bn.Context(0)
rng = np.random.default_rng(11)
G, C, S = 300, 120, 600
X = (rng.random((G, C)) < 0.03) * rng.integers(1, 4, (G, C)).astype(np.float64)
beta = rng.uniform(0.03, 0.2, C)
size = rng.uniform(0.5, 6.0, G)
mu = np.where(np.arange(G) % 3 == 0, 0.5, np.where(np.arange(G) % 3 == 1, 60.0, 4000.0)) * rng.uniform(0.5, 2.0, G)
Mo = oracle.CSC.from_dense(X)
M = bn.Check_input(X, ctx)
d = bn.Main_NB_Bay(M, beta, size, mu, S=S, seed=2024)
nb_size, nb_mu = oracle.post_sample_params(Mo, beta, size, mu)
mean_th = nb_mu + X
var_th = nb_mu + nb_mu ** 2 / nb_size
z = (d.mean(axis=2) - mean_th) / np.sqrt(var_th / S)
kap4 = var_th * (1 + 6 * var_th / nb_size)
vz = (d.var(axis=2, ddof=1) - var_th) / np.sqrt((kap4 + 2 * var_th ** 2) / S)
for cls in range(3):
    sel = np.arange(G) % 3 == cls
    print(cls, 'z max', np.abs(z[sel]).max(), 'z mean', z[sel].mean(), 'vz max', np.abs(vz[sel]).max(), 'vz mean', vz[sel].mean())
i, j = np.unravel_index(np.argmax(np.abs(vz)), vz.shape)
print('worst', i, j, 'cls', i % 3, 'x', X[i, j], 'size', nb_size[i, j], 'mu', nb_mu[i, j], 'var_th', var_th[i, j], 'var', d[i, j].var(ddof=1), 'mean', d[i, j].mean(), 'max draw', d[i, j].max(), 'vz', vz[i, j])
print(np.sort(d[i, j])[-10:])
