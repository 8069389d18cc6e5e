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


def test_header_is_plain_c_and_struct_layout_matches_ctypes(tmp_path):
    """include/clipn.h must compile as C99 and as C++ with nothing but <stdint.h> (no torch / CUDA types in the
    boundary), and the byte layout gcc gives `clipn_gemm_desc` must be the one the ctypes binding uses."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    fields = [f[0] for f in L.GemmDesc._fields_]
    prog = ['#include <stddef.h>', '#include <stdio.h>', '#include "clipn.h"', 'int main(void) {',
            '  printf("sizeof %zu\\n", sizeof(clipn_gemm_desc));']
    prog += [f'  printf("{f} %zu\\n", offsetof(clipn_gemm_desc, {f}));' for f in fields]
    prog += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    assert int(out.pop("sizeof")) == ctypes.sizeof(L.GemmDesc)
    for f in fields:
        assert int(out[f]) == getattr(L.GemmDesc, f).offset, f
    if shutil.which("g++") is not None:
        cpp = tmp_path / "hdr.cpp"
        cpp.write_text('#include "clipn.h"\nint main() { return clipn_version == nullptr; }\n')
        subprocess.run(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I", inc, str(cpp)], check=True)


def test_a_plain_c_program_links_and_calls_the_library(tmp_path):
    """The boundary is usable without Python or torch: a C translation unit that includes only clipn.h links against
    libclipn.so and calls it. Without a GPU a compute entry point must come back with an error code and a message,
    not crash."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "consumer.c"
    src.write_text("""
#include <stdio.h>
#include <string.h>
#include "clipn.h"
int main(void) {
  clipn_gemm_desc d;
  memset(&d, 0, sizeof d);            /* an empty problem: rejected by argument validation before any CUDA call */
  int rc = clipn_gemm(&d, 0);
  printf("version %d rc %d err %s\\n", clipn_version(), rc, clipn_last_error());
  return 0;
}
""")
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:" + os.path.basename(L.LIB_PATH), "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.startswith("version 100 rc -"), out
    assert "gemm" in out


def test_product_path_refuses_cpu_tensors():
    import torch
    from open_clip_b200 import ops
    with pytest.raises(L.ClipnError):
        ops.layernorm_fwd(torch.zeros(4, 64, dtype=torch.bfloat16), torch.ones(64), torch.zeros(64))
