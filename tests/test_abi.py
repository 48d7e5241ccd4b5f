"""CPU checks of the drop-in boundary: libsicp_b200.so loads, exports every function include/sicp_b200.h declares,
the ctypes structs match the C layout, and compute entry points refuse to run (loudly) without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sicp_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sicp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(sicp):
    lib = sicp.lib()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/sicp_b200.h but not exported"
    assert set(sicp.EXPORTS) <= set(names)


def test_struct_layouts_match_the_header(sicp, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sicp_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(sicp_options), sizeof(sicp_result),'
                   ' offsetof(sicp_result, stage_ms), offsetof(sicp_result, pass_pose7), offsetof(sicp_options, gate_d2));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT}/include", str(src), "-o", str(exe)])  # the header is plain C
    so, sr, o_stage, o_pass, o_gate = map(int, subprocess.check_output([str(exe)]).split())
    assert C.sizeof(sicp.Options) == so and C.sizeof(sicp.Result) == sr
    assert sicp.Result.stage_ms.offset == o_stage and sicp.Result.pass_pose7.offset == o_pass and sicp.Options.gate_d2.offset == o_gate


def test_default_options_are_the_reference_constants(sicp):
    o = sicp.default_options(sicp.ALGO_EM)
    assert (o.k_cov, o.epsilon, o.gate_d2, o.min_class_points, o.max_lm_iterations) == (20, 1e-3, 250.0, 400, 400)


def test_no_cpu_fallback(sicp):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    xyz = np.zeros((8, 3), dtype=np.float32)
    with pytest.raises(sicp.SicpError, match="no CUDA device"):
        sicp.Cloud(xyz)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under semantic-icp_b200/ may reference it."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "semantic-icp_b200")):
        if os.sep + "build" in dp or dp.endswith("lib"):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\boracle\b", txt) and re.search(r"import oracle|from oracle|oracle/|sicp_oracle|libsicp_oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
