"""CPU: the C-ABI library loads, exports every symbol include/mtf_b200.h declares, its structs match the ctypes
mirror, the header is plain C, and -- there being no CPU path -- creation fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mtf_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtfb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mtf_b200 import api
    L = api.load_library()
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(api.EXPORTS) == names


def test_header_is_plain_c_and_structs_match(tmp_path):
    from mtf_b200 import api
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mtf_b200.h"\n'
                    'int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mtfb_params), sizeof(mtfb_iter_log), '
                    'offsetof(mtfb_params, epsilon), offsetof(mtfb_iter_log, rejected), sizeof(mtfb_pf_params), '
                    'offsetof(mtfb_pf_params, seed), sizeof(mtfb_est_params), offsetof(mtfb_est_params, seed)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(prog), "-o", str(exe)])
    a, b, c, d, e, f, g, h = map(int, subprocess.check_output([str(exe)]).split())
    assert a == C.sizeof(api.Params) and b == C.sizeof(api.IterLog)
    assert c == api.Params.epsilon.offset and d == api.IterLog.rejected.offset
    assert e == C.sizeof(api.PFParams) and f == api.PFParams.seed.offset
    assert g == C.sizeof(api.EstParams) and h == api.EstParams.seed.offset


def test_default_params_and_loud_failure_without_gpu():
    import torch
    from mtf_b200 import api
    p = api.default_params()
    assert (p.resx, p.resy, p.max_iters, p.chained_warp, p.grad_eps) == (50, 50, 30, 1, 1e-8)
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the no-device error path cannot be exercised")
    with pytest.raises(api.MTFError) as e:
        api.BatchTracker(p)
    assert e.value.status == 5 and "no CPU path" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under mtf_b200/ or include/ may import, include or link it"""
    bad = []
    for base in ("mtf_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_obj" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle_lib|mtf_oracle|libmtf_oracle|from oracle|import oracle", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    out = subprocess.check_output(["ldd", os.path.join(ROOT, "mtf_b200", "libmtf_b200.so")], text=True)
    assert "oracle" not in out
