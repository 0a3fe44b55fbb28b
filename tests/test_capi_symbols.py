"""CPU checks of the drop-in boundary: the library builds, loads, and exports every symbol include/azb.h declares;
host-only entry points behave.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "azb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(azb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(capi):
    syms = _declared_symbols()
    assert len(syms) >= 30
    L = C.CDLL(capi.lib_path())
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/azb.h but not exported"
    assert sorted(capi.SIGNATURES) == syms  # the ctypes binding covers the whole header, nothing else


def test_version_and_strerror(capi):
    L = capi.lib()
    assert L.azb_version() == 100
    assert L.azb_strerror(0) == b"ok"
    assert b"capacity" in L.azb_strerror(capi.ERR_CAPACITY)


def test_config_default_matches_the_example(capi):
    cfg = capi.default_config(19, 512)  # graph-state/examples/04-c21-tree.rs:33-54,133-138
    assert cfg.struct_size == C.sizeof(capi.Config)
    assert list(cfg.n_as_tol)[:3] == [200, 50, 50] and cfg.n_as_tol_len == 3 and cfg.n_as_tol_default == 25
    assert list(cfg.mlp_hidden) == [512, 1024, 512] and cfg.max_steps == 800 and cfg.c_lower == 2.0
    with pytest.raises(capi.AzbError):
        capi.default_config(4, 1)
    with pytest.raises(capi.AzbError):
        capi.default_config(65, 1)


def test_root_generator_is_the_oracles(capi, orc):
    for n, seed, first in ((19, 0, 0), (64, 5, 1000), (7, 2, 3)):
        p, m = capi.generate_roots(seed, first, 64, n)
        po, mo = orc.generate_roots(seed, first, 64, n)
        assert np.array_equal(p, po) and np.array_equal(m, mo)


def test_create_without_gpu_fails_cleanly(capi):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = capi.default_config(19, 8)
    with pytest.raises(capi.AzbError) as e:
        capi.Handle(cfg)
    assert e.value.code == capi.ERR_CUDA  # no silent CPU fallback


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "azdopt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libazb_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
