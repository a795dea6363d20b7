"""CPU tests of the multi-GPU host logic (world_size 2, gloo): gene-block partitioning, CSC row
splitting and the all-reduce hook the C library calls for its cross-shard steps.  The hook is
exercised through the same ctypes function-pointer type the library uses, on host buffers."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from util import synth_counts


def test_balanced_gene_blocks():
    from baynorm_b200 import dist
    rng = np.random.default_rng(0)
    nnz = (rng.pareto(1.2, 5000) * 50).astype(np.int64)
    for world in (1, 2, 4, 8):
        b = dist.balanced_gene_blocks(nnz, world, n_cells=1000)
        assert b[0] == 0 and b[-1] == 5000 and (np.diff(b) > 0).all() and len(b) == world + 1
        cost = np.add.reduceat(1000 + 4.0 * nnz, b[:-1])
        assert cost.max() / cost.mean() < 1.25
    # the weight of a stored entry against a cell is the caller's (4: first-generation fit, ~1.5: fit version 2 with S = 20 draws)
    heavy = np.r_[np.full(10, 10000.0), np.zeros(990)]
    b4, b0 = dist.balanced_gene_blocks(heavy, 2, n_cells=100, nnz_weight=4.0), dist.balanced_gene_blocks(heavy, 2, n_cells=100, nnz_weight=0.0)
    assert b0[1] == 500 and b4[1] < 20                 # weight 0: equal gene counts; weight 4: the ten heavy genes are half the cost
    b = dist.balanced_gene_blocks(np.zeros(3), 3)
    assert list(b) == [0, 1, 2, 3]


def test_split_csc_rows(oracle):
    from baynorm_b200 import dist
    X, *_ = synth_counts(97, 41, seed=4, m0=1.0)
    M = oracle.CSC.from_dense(X)
    parts = []
    for g0, g1 in ((0, 30), (30, 31), (31, 97)):
        G, Cn, p, i, x = dist.split_csc_rows(M.G, M.C, M.p, M.i, M.x, g0, g1)
        S = oracle.CSC(G, Cn, p, i, x)
        assert np.array_equal(S.to_dense(), X[g0:g1])
        ref = M.select_rows(g0, g1)
        assert np.array_equal(ref.p, p) and np.array_equal(ref.i, i)
        parts.append(S.to_dense())
    assert np.array_equal(np.vstack(parts), X)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as td
    from baynorm_b200 import _lib, dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hook = dist.make_allreduce_hook(None, None)

        def tramp(user, buf, count, dtype, op, stream):
            return hook(buf, count, dtype, op, stream)
        cb = _lib.ALLREDUCE_FN(tramp)            # exactly what bn_ctx_set_allreduce receives
        out = {}
        # per-cell totals of a gene-sharded matrix -> BETA (sum), then min/max and OLS sums
        X, mu, size, beta = synth_counts(60, 50, seed=21, m0=1.0)
        bounds = dist.balanced_gene_blocks((X > 0).sum(axis=1), world, n_cells=50)
        Xs = X[bounds[rank]:bounds[rank + 1]]
        col = np.ascontiguousarray(Xs.sum(axis=0))
        assert cb(None, col.ctypes.data, col.size, 0, 0, None) == 0
        out["colsum"] = col
        mm = np.array([Xs.max(), -Xs.max()], dtype=np.float64) * (rank + 1)
        assert cb(None, mm.ctypes.data, 2, 0, 1, None) == 0
        out["min"] = mm
        mx = np.array([float(rank)], dtype=np.float64)
        assert cb(None, mx.ctypes.data, 1, 0, 2, None) == 0
        out["max"] = mx
        assert cb(None, mx.ctypes.data, 1, 1, 0, None) == 2       # unsupported dtype is reported, not ignored
        out["full_colsum"] = X.sum(axis=0)
        q.put((rank, out))
    finally:
        td.destroy_process_group()


def test_allreduce_hook_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert np.array_equal(res[r]["colsum"], res[r]["full_colsum"])      # integer counts: exact
        assert res[r]["max"][0] == 1.0
    assert np.array_equal(res[0]["min"], res[1]["min"])


def test_sharded_prior_statistics_match_unsharded(oracle):
    """The three cross-shard reductions of Prior_fun reproduce the unsharded numbers: NaN fill /
    bounds from min-max of MME sizes, and AdjustSIZE_fun from min/max of BB_SIZE + five sums."""
    X, mu, size, beta = synth_counts(200, 120, seed=8, m0=0.5)
    M = oracle.CSC.from_dense(X)
    full = oracle.Prior_fun(M, beta)
    from baynorm_b200 import dist
    bounds = dist.balanced_gene_blocks((X > 0).sum(axis=1), 3, n_cells=120)
    shards = [M.select_rows(bounds[k], bounds[k + 1]) for k in range(3)]
    pri = [oracle.EstPrior(s, beta) for s in shards]
    lo = min(np.nanmin(p["SIZE"]) for p in pri)
    hi = max(np.nanmax(p["SIZE"]) for p in pri)
    sizes = [np.where(np.isnan(p["SIZE"]), lo, p["SIZE"]) for p in pri]
    assert np.array_equal(np.concatenate(sizes), full["MME_prior"]["MME_SIZE"])
    bbs = [oracle.BB_Fun(s, beta, p["MU"], sz, lo, np.ceil(hi)) for s, p, sz in zip(shards, pri, sizes)]
    assert np.array_equal(np.concatenate(bbs), full["BB_prior"]["BB_SIZE"])
    bmin, bmax = min(b.min() for b in bbs), max(b.max() for b in bbs)
    sums = np.zeros(5)
    for b, sz in zip(bbs, sizes):
        fit = (b < bmax) & (b > bmin)
        x, y = np.log(sz[fit]), np.log(b[fit])
        sums += [fit.sum(), x.sum(), y.sum(), (x * x).sum(), (x * y).sum()]
    n, sx, sy, sxx, sxy = sums
    slope = (sxy - sx * sy / n) / (sxx - sx * sx / n)
    icpt = sy / n - slope * sx / n
    adj = np.concatenate([np.exp(icpt + slope * np.log(sz)) for sz in sizes])
    assert np.abs(adj / full["MME_SIZE_adjust"] - 1).max() < 1e-11
