"""Timings of the path's smaller entry points on the config-2 matrix (20k x 10k, ~50 % dense) and the config-7
matrix (33k x 20k, ~8 %): DownSampling, default BETA, BetaFun, SyntheticControl, MME alone, gene-major build.
Device-resident inputs/outputs, CUDA events on the library's stream, best of 5."""
import ctypes as CT
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import baynorm_b200 as bn
from baynorm_b200.api import _ptr
from bench import SEED, synth_params


def timed(ctx, ext, fn, reps=5):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        fn()
        e1.record(ext)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ctx = bn.Context(0)
    dev = torch.device("cuda", 0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    lib = ctx.lib
    out = {}
    for name, (G, C, m0) in {"cfg2": (20000, 10000, 3.013), "cfg7": (33000, 20000, -0.675)}.items():
        mu, size, beta = synth_params(G, C, m0)
        M = bn.DeviceMatrix.synth_nb(ctx, mu, size, beta, seed=SEED, gene0=0)
        f64 = dict(dtype=torch.float64, device=dev)
        d_beta = torch.from_numpy(beta).to(dev)
        d_vecC = torch.empty(C, **f64)
        d_G = [torch.empty(G, **f64) for _ in range(3)]
        d_sel = torch.empty(G, dtype=torch.int32, device=dev)
        d_stats = torch.empty(4 * G, **f64)
        r = {"G": G, "C": C, "nnz": int(M.nnz)}
        t = timed(ctx, ext, lambda: ctx.check(lib.bn_beta_default(ctx.h, M.h, 0.06, _ptr(d_vecC))))
        r["beta_default_ms"] = t
        r["beta_default_GBs"] = M.nnz * 8 / t / 1e6
        def build():
            lib.bn_mat_drop_gene_major(M.h)
            ctx.check(lib.bn_mme(ctx.h, M.h, _ptr(d_beta), _ptr(d_G[0]), _ptr(d_G[1]), _ptr(d_G[2])))
        t_all = timed(ctx, ext, build)
        t_mme = timed(ctx, ext, lambda: ctx.check(lib.bn_mme(ctx.h, M.h, _ptr(d_beta), _ptr(d_G[0]), _ptr(d_G[1]), _ptr(d_G[2]))))
        r["gene_major_build_ms"] = t_all - t_mme
        r["mme_ms"] = t_mme
        r["mme_GBs"] = M.nnz * 16 / t_mme / 1e6
        t = timed(ctx, ext, lambda: ctx.check(lib.bn_beta_fun(ctx.h, M.h, 0.06, _ptr(d_vecC), _ptr(d_sel), _ptr(d_stats))))
        r["beta_fun_ms"] = t
        r["beta_fun_selected"] = int(d_sel.sum().item())
        if G * C * 8 <= 4 << 30:
            d_dense = torch.empty(G * C, **f64)
            t = timed(ctx, ext, lambda: ctx.check(lib.bn_synthetic_control(ctx.h, M.h, _ptr(d_beta), 7, _ptr(d_G[0]), _ptr(d_dense))))
            r["synthetic_control_ms"] = t
            r["synthetic_control_GBs"] = G * C * 8 / t / 1e6
            d_out = torch.empty(G * C, **f64)
            t = timed(ctx, ext, lambda: ctx.check(lib.bn_downsample(ctx.h, G, C, _ptr(d_dense), _ptr(d_beta), 9, _ptr(d_out))))
            r["downsample_ms"] = t
            r["downsample_GBs"] = G * C * 16 / t / 1e6
            del d_dense, d_out
        out[name] = r
        M.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
