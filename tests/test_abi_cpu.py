"""CPU checks of the boundary: the C-ABI library builds/loads, exports every symbol include/sgr.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side helpers behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sgr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from surge_b200 import native as N

    lib = N.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sgr.h but not exported by libsgr.so"
    bound = {n for n, _, _ in N.ABI}
    assert set(declared) <= bound, sorted(set(declared) - bound)
    assert lib.sgr_abi_version() == 1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from surge_b200 import ReplayEngine, SgrError
    from surge_b200 import native as N

    with pytest.raises(SgrError) as ei:
        ReplayEngine(0)
    assert ei.value.code == N.SGR_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "surge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "sgr_oracle" not in text and "liborc" not in text, f


def test_program_structs_match_the_header_sizes():
    from surge_b200 import native as N
    from surge_b200 import programs as P

    assert C.sizeof(N.sgr_op) == 8 and C.sizeof(N.sgr_rule) == 8 + 8 * N.MAX_OPS
    assert C.sizeof(N.sgr_fold_program) == 16 + 16 + C.sizeof(N.sgr_rule) * N.MAX_TYPES
    p = P.counter_program()
    assert p.state_bytes == 16 and p.n_types == 4 and p.rules[0].n_ops == 2 and p.rules[3].exists_rule == N.THROW
    b = P.bank_account_program()
    assert b.state_bytes == 64 and b.n_f64_fields == 1 and b.f64_field_off[0] == 16


def test_formats_roundtrip():
    from surge_b200 import formats as F

    r = F.counter_records([0, 1], [1, 2], [5, 5], [7, -3])
    assert r.view(np.uint8).size == 128 and int(r["arg0"][1]) == -3
    rec = F.bank_created_record(3, 1, "00000000-0000-0000-0000-000000001234", "Jane Doe", "1234", 1000.0)
    assert len(rec) == 64
    b, off = F.pack_var_records([0, 2], [1, 2], [0, 0], [b"\x01\0\0\0" + bytes(29), b""])
    assert list(off) == [0, 16 + 48, 16 + 48 + 16] and len(b) == 80
    assert F.counter_state_json("a", 4, 4) == b'{"aggregateId":"a","count":4,"version":4}'
