#!/usr/bin/env python
"""Freeze the oracle's outputs on the bundled fixture (BASELINE.json configs[0]).

These are NOT reference outputs (the reference cannot run in this image: no R); they pin the
oracle restatement itself, so that later edits to oracle/ cannot silently move the target the
CUDA path is compared with.  Run from the repo root: python tests/golden/make_oracle_golden.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import oracle as O  # noqa: E402

fx = O.load_fixture()
A, beta = fx["inputdata"], fx["inputbeta"]
M = O.CSC.from_dense(A)
r = O.bayNorm(M, BETA_vec=beta, mean_version=True)
out = {
    "MME_MU": r["PRIORS"]["MME_prior"]["MME_MU"],
    "MME_SIZE": r["PRIORS"]["MME_prior"]["MME_SIZE"],
    "BB_SIZE": r["PRIORS"]["BB_prior"]["BB_SIZE"],
    "MME_SIZE_adjust": r["PRIORS"]["MME_SIZE_adjust"],
    "Bay_out_mean": r["Bay_out"],
    "mean_bundled_priors": O.Main_mean_NB_Bay(M, beta, fx["size"], fx["mu"]),
    "mode_bundled_priors": O.Main_mode_NB_Bay(M, beta, fx["size"], fx["mu"]),
}
np.savez_compressed(Path(__file__).with_name("oracle_fixture_outputs.npz"), **out)
for k, v in out.items():
    print(k, v.shape, float(np.asarray(v).sum()))
