"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/baynorm_b200.h declares, the ctypes table covers exactly those symbols, and the product
fails loudly (never falls back to a CPU path) when no GPU is present."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "baynorm_b200.h"


def declared_symbols():
    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"BN_API\s+[\w\s\*]+?\b(bn_\w+)\s*\(", txt)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("bn_ctx_create", "bn_mat_from_csc", "bn_beta_default", "bn_mme", "bn_fit_size", "bn_adjust_size",
                 "bn_prior", "bn_post_mean", "bn_post_mode", "bn_post_sample", "bn_downsample", "bn_marginal_nb_1d",
                 "bn_marginal_nb_2d", "bn_gradient_nb_1d", "bn_gradient_nb_2d", "bn_row_means", "bn_row_vars",
                 "bn_mme_dense", "bn_ctx_set_allreduce"):
        assert must in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol():
    from baynorm_b200 import _lib
    assert _lib.LIB_PATH.exists(), "build with __graft_entry__.build() / make -C baynorm_b200/csrc"
    lib = C.CDLL(str(_lib.LIB_PATH))
    for s in declared_symbols():
        assert hasattr(lib, s), "libbaynorm_b200.so does not export %s" % s
    assert lib.bn_version() == 100


def test_ctypes_table_matches_header():
    from baynorm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()


def test_no_cpu_fallback_without_a_gpu():
    """On a box without a GPU the product must raise, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import baynorm_b200 as bn
    with pytest.raises(bn.BayNormError) as e:
        bn.Context(0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)
    import numpy as np
    with pytest.raises(bn.BayNormError):
        bn.Main_mean_NB_Bay(np.ones((3, 4)), np.full(4, 0.1), np.ones(3), np.ones(3))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under baynorm_b200/ may include, import, link or load it."""
    for f in (ROOT / "baynorm_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h") or f.name == "Makefile":
            t = f.read_text()
            assert not re.search(r"#\s*include\s*[\"<][^\">]*oracle", t), f
            assert not re.search(r"^\s*(from|import)\s+oracle", t, flags=re.M), f
            assert "libbaynorm_oracle" not in t and "baynorm_oracle.c" not in t.replace("oracle/baynorm_oracle.c", ""), f


def test_argument_errors_are_raised_before_any_gpu_work():
    import baynorm_b200 as bn
    with pytest.raises(ValueError, match="Only one of mode_version"):
        bn.api._myFunc(True, True, False)
    assert bn.api._myFunc(False, False, False) is bn.Main_NB_Bay
    assert bn.api._myFunc(True, False, False) is bn.Main_mode_NB_Bay
    assert bn.api._myFunc(False, True, False) is bn.Main_mean_NB_Bay
    assert bn.api._myFunc(False, True, True) is bn.Main_mean_NB_spBay
    for name in ("BetaFun", "SyntheticControl", "BB_Fun", "Prior_fun", "bayNorm", "bayNorm_sup", "DownSampling"):
        assert callable(getattr(bn, name)), name            # the reference's exported functions on this path


def test_r_shim_matches_the_c_abi():
    """r_shim/baynorm_b200_shim.c cannot be built here (no R), but it can be type-checked: declarations-only stand-ins
    for R.h / Rinternals.h (tests/rmock/) plus the real include/baynorm_b200.h.  Any call of a bn_* entry point with
    the wrong arity or pointer types, or a registered routine with the wrong argument count, fails here."""
    import subprocess
    src = ROOT / "r_shim" / "baynorm_b200_shim.c"
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                        "-Werror=int-conversion", "-I", str(ROOT / "tests" / "rmock"), "-I", str(ROOT / "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = src.read_text()
    # every registered routine: the arity in CallEntries equals the number of SEXP parameters of its definition
    for name, arity in re.findall(r'\{"(\w+)", \(DL_FUNC\)&\w+, (\d+)\}', text):
        m = re.search(r"SEXP %s\(([^)]*)\)" % name, text)
        assert m, name
        assert m.group(1).count("SEXP") == int(arity), (name, arity, m.group(1))
    # ALL 15 names the reference registers keep their arities (src/RcppExports.cpp:211-228)
    want = {"_bayNorm_DownSampling": 2, "_bayNorm_MarginalF_NB_1D": 4, "_bayNorm_MarginalF_NB_2D": 3, "_bayNorm_GradientFun_NB_1D": 4,
            "_bayNorm_GradientFun_NB_2D": 3, "_bayNorm_Main_NB_Bay": 6, "_bayNorm_Main_mean_NB_Bay": 6, "_bayNorm_Main_mode_NB_Bay": 6,
            "_bayNorm_rowMeansFast": 1, "_bayNorm_rowVarsFast": 2, "_bayNorm_EstPrior_sprcpp": 1, "_bayNorm_EstPrior_rcpp": 1,
            "_bayNorm_t_sp": 1, "_bayNorm_asMatrix": 5, "_bayNorm_Main_mean_NB_spBay": 6}
    assert len(want) == 15
    got = dict((n, int(a)) for n, a in re.findall(r'\{"(\w+)", \(DL_FUNC\)&\w+, (\d+)\}', text))
    for n, a in want.items():
        assert got.get(n) == a, (n, got.get(n))
