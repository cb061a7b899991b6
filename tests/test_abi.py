"""The C-ABI shared library builds for sm_100a on a CPU box, loads, and exports every symbol the header declares.
No compute calls here (no GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "emrt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emrt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_built):
    from emrt_b200 import _lib
    lib = ctypes.CDLL(lib_built)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/emrt_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert _lib.load().emrt_version() == 1


def test_library_is_sm100a_only(lib_built):
    out = subprocess.run(["cuobjdump", "-lelf", lib_built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_on_cpu_tensors(lib_built):
    import torch
    import emrt_b200
    from emrt_b200 import ops
    with pytest.raises(emrt_b200.EmrtError):
        ops.upsample2x(torch.zeros(1, 1, 2, 2))
    m = emrt_b200.MSDeformableAttention(64, 2, 3, 6)
    with pytest.raises(emrt_b200.EmrtError):
        m(torch.zeros(1, 84, 64), torch.zeros(1, 84, 3, 2), torch.zeros(1, 84, 64), [(8, 8), (4, 4), (2, 2)])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "emrt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f"{f} imports the oracle"


def test_state_dict_keys_match_reference_layout():
    import emrt_b200
    m = emrt_b200.MSDeformableAttention(256, 8, 3, 6)
    sd = m.state_dict()
    assert tuple(sd["sampling_offsets.weight"].shape) == (256, 288)      # Paddle Linear weight is [in, out]
    assert tuple(sd["attention_weights.weight"].shape) == (256, 144)
    assert tuple(sd["value_proj.weight"].shape) == (256, 256)
    assert tuple(sd["output_proj.weight"].shape) == (256, 256)
    assert set(k.split(".")[0] for k in sd) == {"sampling_offsets", "attention_weights", "value_proj", "output_proj"}
    import oracle as O
    assert abs(sd["sampling_offsets.bias"].numpy() - O.msda_reset_parameters(256, 8, 3, 6)).max() < 1e-6


def test_paddle_shim_parses_and_refuses_to_import_without_paddle():
    """emrt_b200/paddle_shim.py cannot run here (no PaddlePaddle); it must at least be valid Python, bind only symbols the
    C ABI exports, and fail with a clear ImportError instead of degrading to anything else."""
    import ast
    import importlib
    import re
    path = os.path.join(ROOT, "emrt_b200", "paddle_shim.py")
    src = open(path).read()
    ast.parse(src)
    from emrt_b200 import _lib
    for name in set(re.findall(r"\.(emrt_[a-z0-9_]+)\(", src)):
        assert name in _lib.SIGNATURES, name
    try:
        import paddle  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="PaddlePaddle"):
            importlib.import_module("emrt_b200.paddle_shim")
