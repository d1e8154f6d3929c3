"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/apop_b200.h declares; struct layouts match the reference's; and without a
GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "apop_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apop_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_typed():
    from apophenia_b200 import abi
    lib = abi.load_library()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/apop_b200.h but not exported"
        assert n in abi.SIGNATURES, f"{n} has no ctypes signature"
    out = subprocess.check_output(["nm", "-D", "--defined-only", abi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (\w+)", out))
    assert set(names) <= exported
    assert all(e.startswith("apop_b200_") for e in exported), "only the C-ABI is exported"


def test_struct_layouts_match_reference():
    from apophenia_b200 import abi
    # gsl_vector {size, stride, data, block, owner}; gsl_matrix {size1, size2, tda, data, block, owner}
    assert [f[0] for f in abi.gsl_vector._fields_] == ["size", "stride", "data", "block", "owner"]
    assert [f[0] for f in abi.gsl_matrix._fields_] == ["size1", "size2", "tda", "data", "block", "owner"]
    # apop_data field order (apop.m4.h:72-81) and apop_model (apop.m4.h:98-115)
    assert [f[0] for f in abi.apop_data._fields_] == ["vector", "matrix", "names", "text", "textsize", "weights", "more", "error"]
    assert [f[0] for f in abi.apop_model._fields_][:8] == ["name", "vsize", "msize1", "msize2", "dsize", "data", "parameters", "info"]
    assert [f[0] for f in abi.apop_model._fields_][8:15] == ["estimate", "p", "log_likelihood", "cdf", "constraint", "draw", "prep"]
    assert abi.apop_model.name.size == 101 and abi.apop_settings_type.name.size == 101


def test_data_views_keep_strides():
    from apophenia_b200 import abi
    big = np.arange(40.0).reshape(5, 8)
    d = abi.Data(matrix=big[:, 2:6], vector=np.arange(10.0)[::2])
    assert d._m.size1 == 5 and d._m.size2 == 4 and d._m.tda == 8
    assert d._v.size == 5 and d._v.stride == 2
    assert d.struct.matrix.contents.data[8] == big[1, 2]
    page = d.add_page(abi.Data(vector=[0.0, 1.0], title="<categories for vector>"))
    assert d.struct.more.contents.names.contents.title == b"<categories for vector>"
    assert page.struct.vector.contents.size == 2


def test_no_cpu_fallback():
    """without a B200 every entry fails loudly: init reports, LL is NaN, score is NaN"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    import apophenia_b200 as ab
    with pytest.raises(ab.B200Error):
        ab.init(0)
    assert "no CPU fallback" in ab.last_error()
    X = np.ones((4, 2)); y = np.zeros(4)
    d = ab.Data(matrix=X, vector=y)
    m = ab.probit_model([0.1, 0.2], d)
    assert np.isnan(ab.log_likelihood(ab.PROBIT, d, m))
    assert np.all(np.isnan(ab.score(ab.PROBIT, d, m)))
    with pytest.raises(ab.B200Error):
        ab.pin(d)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under apophenia_b200/ may reference it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "apophenia_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
