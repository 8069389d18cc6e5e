"""CPU suite: libclipn.so loads and exports every symbol include/clipn.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from open_clip_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "clipn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clipn_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built():
    assert os.path.exists(L.LIB_PATH), "run `python -m open_clip_b200.build` (or __graft_entry__.build())"


def test_every_header_symbol_is_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 25
    handle = ctypes.CDLL(L.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/clipn.h but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes signature in open_clip_b200/_lib.py"
    assert set(L.SIGNATURES) == set(syms)


def test_version_and_error_string():
    lib = L.lib()
    assert lib.clipn_version() == 100
    assert isinstance(lib.clipn_last_error(), bytes)


def test_gemm_desc_layout_matches_header():
    """ctypes struct must mirror `struct clipn_gemm_desc` field-for-field (names and order)."""
    text = open(os.path.join(ROOT, "include", "clipn.h")).read()
    body = re.search(r"typedef struct clipn_gemm_desc \{(.*?)\} clipn_gemm_desc;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"([a-z0-9_]+)\s*$", part.strip())[0])
    assert names == [f[0] for f in L.GemmDesc._fields_]


def test_product_path_refuses_cpu_tensors():
    import torch
    from open_clip_b200 import ops
    with pytest.raises(L.ClipnError):
        ops.layernorm_fwd(torch.zeros(4, 64, dtype=torch.bfloat16), torch.ones(64), torch.zeros(64))
