"""A/B of the 3-D sampler's side streams on the config-6 matrix: posterior draws alone, serial vs pipelined."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import baynorm_b200 as bn
from baynorm_b200.api import _ptr
from bench import SEED, synth_params, CONFIGS

cfgn = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cfg = CONFIGS[cfgn]
G, C, S = cfg["G"], cfg["C"], cfg["S"]
ctx = bn.Context(0)
dev = torch.device("cuda", 0)
mu, size, beta = synth_params(G, C, cfg["m0"])
M = bn.DeviceMatrix.synth_nb(ctx, mu, size, beta, seed=SEED, gene0=0)
d_beta, d_size, d_mu = (torch.from_numpy(a).to(dev) for a in (beta, size, mu))
tile = min(C, int(8 * 2**30 // (G * S * 8)))
out = torch.empty(G * tile * S, dtype=torch.float64, device=dev)
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
def run():
    for c0 in range(0, C, tile):
        nc = min(tile, C - c0)
        ctx.check(ctx.lib.bn_post_sample(ctx.h, M.h, _ptr(d_beta), _ptr(d_size), _ptr(d_mu), c0, nc, S, SEED, _ptr(out)))
for mode in (1, 0, 1, 0):
    ctx.lib.bn_debug_tune(ctx.h, 3, mode)
    run(); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(3): run()
    e1.record(ext); e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("serial" if mode else "pipelined", "%.2f ms  %.0f GB/s" % (ms, G * C * S * 8 / ms / 1e6), flush=True)
