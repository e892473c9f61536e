"""The C-ABI library: loads, exports every symbol include/ani_b200.h declares, struct layouts
match the ctypes mirror, and argument validation works without a GPU (no compute calls)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from helpers import ROOT

HEADER = os.path.join(ROOT, "include", "ani_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"\b(ani_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from torchani_b200 import build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = C.CDLL(lib_path)
    declared = _declared_symbols()
    assert len(declared) >= 11
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ani_b200.h but not exported"
    from torchani_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes prototypes out of sync with the header"
    assert _lib.lib().ani_b200_abi_version() == 4


def test_struct_layouts_match_header(tmp_path):
    from torchani_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ani_b200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ani_aev_params), sizeof(ani_grid),'
        ' sizeof(ani_mlp_species), sizeof(ani_mlp_model), offsetof(ani_grid, n_real),'
        ' offsetof(ani_mlp_model, sp), offsetof(ani_aev_params, cos_z));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(v) for v in out]
    assert sizes == [C.sizeof(_lib.AEVParams), C.sizeof(_lib.Grid), C.sizeof(_lib.MLPSpecies),
                     C.sizeof(_lib.MLPModel), _lib.Grid.n_real.offset, _lib.MLPModel.sp.offset,
                     _lib.AEVParams.cos_z.offset]


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call is made."""
    from torchani_b200 import _lib
    L = _lib.lib()
    assert L.ani_b200_error_string(0) == b"ok"
    assert b"argument" in L.ani_b200_error_string(-1)
    # null pointers
    assert L.ani_b200_build_cells(None, None, 1, 1, None, 0, 0, 5.1, 64, None, None, None, None, None, None,
                                  None, None, None, None) == -1
    assert L.ani_b200_aev_forward(None, None, None, None, None, None, None, None, 0, 0, 0, None, None, 0, 0, None, None,
                                  128, None, None) == -1
    # unsupported AEV configuration (ShfA x ShfZ must be 8x4 or 4x8) is reported as such
    p = _lib.AEVParams()
    p.num_species, p.n_shf_r, p.n_shf_a, p.n_shf_z = 7, 16, 5, 5
    p.rcr, p.rca, p.zeta = 5.1, 3.5, 14.1
    assert L.ani_b200_aev_forward(C.byref(p), None, None, None, None, None, None, None, 0, 0, 0, None, None, 0, 0, None,
                                  None, 128, None, None) == -2
    with pytest.raises(_lib.ANIB200Error):
        _lib.check(-2, "x")


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under torchani_b200/ may reference it."""
    pkg = os.path.join(ROOT, "torchani_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle." not in text, \
                    f"{f} uses the oracle"
    code = "import sys; import torchani_b200.models, torchani_b200.parallel; print(any(m.startswith('oracle') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, check=True)
    assert out.stdout.strip() == "False"
