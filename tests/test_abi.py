"""CPU-only: the C-ABI library loads and exports every symbol the header declares
(no compute call is made: there is no GPU in the build container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "barefoot_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from barefoot_framework_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    assert sorted(_lib.SIGNATURES) == names


def header_arg_counts():
    """name -> number of parameters of every prototype in the header."""
    src = open(os.path.join(ROOT, "include", "barefoot_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(bf_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_binding_and_integration_stub_have_the_header_arity():
    """Every ctypes signature has as many arguments as its prototype, and the stub INTEGRATION.md shows a
    reference maintainer binds the sweep and the training-side pre-scaling with the same argument lists."""
    import ast
    from barefoot_framework_b200 import _lib
    counts = header_arg_counts()
    for name, (_, args) in _lib.SIGNATURES.items():
        assert counts[name] == len(args), (name, counts[name], len(args))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in ("bf_gp_posterior", "bf_prescale_train"):
        m = re.search(r"lib\.%s\.argtypes = (\[[^\]]*\])" % name, doc)
        assert m, name
        listed = len(ast.parse(m.group(1), mode="eval").body.elts)
        assert listed == counts[name], (name, listed, counts[name])


def test_library_exports_every_symbol():
    from barefoot_framework_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.bf_version() == 100
    # pure host arithmetic helpers
    assert lib.bf_padded_n(500) == 512 and lib.bf_padded_n(32) == 32 and lib.bf_padded_n(1) == 32
    nb = 16
    assert lib.bf_linv_packed_doubles(500) == 1024 * nb * (nb + 1) // 2
    assert lib.bf_posterior_tile(500) == 24
    assert 148 <= lib.bf_posterior_parts(500, 10 ** 6, 1) <= (10 ** 6 + 23) // 24
    assert lib.bf_posterior_parts(500, 10 ** 5, 1000) >= 1
    assert lib.bf_posterior_tile(5000) == 0


def test_product_has_no_cpu_fallback():
    """The product must fail loudly without CUDA rather than route through the oracle."""
    import torch
    from barefoot_framework_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.BarefootCudaError):
        _lib.require_cuda()
    from barefoot_framework_b200 import engine
    with pytest.raises(_lib.BarefootCudaError):
        engine.reification([[1.0, 2.0]], [[1.0, 1.0]])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "barefoot_framework_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_set_chunks_fill_whole_waves():
    """engine._whole_waves: a memory-limited chunk of a chunked set-up is rounded down to a multiple of the SM
    count (the batched Cholesky / inverse / solve kernels run one CTA per set and SM), unless it already holds
    every set or is smaller than one wave."""
    from barefoot_framework_b200 import engine
    saved = engine._sm_count
    engine._sm_count = 148                       # B200; no device query on the CPU
    try:
        assert engine._whole_waves(963, 150000) == 888
        assert engine._whole_waves(476, 1000) == 444
        assert engine._whole_waves(148, 1000) == 148
        assert engine._whole_waves(100, 1000) == 100          # less than one wave: left alone
        assert engine._whole_waves(963, 963) == 963           # one chunk holds every set
        assert engine._whole_waves(963, 500) == 963
    finally:
        engine._sm_count = saved


def test_packed_operand_cells_are_all_written_or_cleared():
    """The packed L^-1 operand is allocated uninitialised.  bf_factor_finish must leave no cell of it undefined
    that the sweep can read: the inverse kernels write the n x n part shifted by np - n (bf_pack_store: natural
    lower block triangle incl. the full natural diagonal blocks), bf_packed_clear_kernel zeroes the shifted
    diagonal blocks and the leading ceil((np - n) / 4) k-groups of every later row block first.  This restates
    those index rules and checks that together they cover every cell of the block-lower-triangular storage."""
    import numpy as np
    for n in list(range(1, 200)) + [333, 480, 497, 500, 511, 512, 513, 543, 544, 700]:
        np_ = (n + 31) // 32 * 32
        sh, nb = np_ - n, np_ // 32
        cover = np.zeros((np_, np_), dtype=bool)
        lead = (sh + 3) // 4
        for b in range(nb):
            cover[32 * b:32 * b + 32, :4 * min(lead, 8 * b)] = True
            cover[32 * b:32 * b + 32, 32 * b:32 * b + 32] = True
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        m = (j // 32) <= (i // 32)
        cover[i[m] + sh, j[m] + sh] = True
        ip, jp = np.meshgrid(np.arange(np_), np.arange(np_), indexing="ij")
        stored = jp < 32 * (ip // 32 + 1)
        assert (cover | ~stored).all(), n
