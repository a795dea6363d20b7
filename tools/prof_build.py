"""Gene-major build alone on the config-2 and config-7 matrices (for ncu launch lists)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import baynorm_b200 as bn
from baynorm_b200.api import _ptr
from bench import SEED, synth_params

ctx = bn.Context(0)
dev = torch.device("cuda", 0)
for name, (G, C, m0) in {"cfg2": (20000, 10000, 3.013), "cfg7": (33000, 20000, -0.675)}.items():
    mu, size, beta = synth_params(G, C, m0)
    M = bn.DeviceMatrix.synth_nb(ctx, mu, size, beta, seed=SEED, gene0=0)
    d = [torch.empty(G, dtype=torch.float64, device=dev) for _ in range(3)]
    d_beta = torch.from_numpy(beta).to(dev)
    for _ in range(3):
        ctx.lib.bn_mat_drop_gene_major(M.h)
        ctx.check(ctx.lib.bn_mme(ctx.h, M.h, _ptr(d_beta), _ptr(d[0]), _ptr(d[1]), _ptr(d[2])))
    ctx.synchronize()
    M.close()
