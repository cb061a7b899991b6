"""Imports the reference's OWN hot-path sources from /root/reference and runs them on oracle/paddle_on_torch.py.
TEST INFRASTRUCTURE ONLY — used here (the build container) to generate the golden vectors under tests/golden/ and
to cross-check oracle/emrt_oracle.py; /root/reference does not exist on the GPU box, so nothing at run time there
imports this module (``available()`` is False and the tests that need it skip).

No reference source is copied: modules are loaded in place by path.  The reference's package ``__init__`` files
import every backbone / dataset of the repo (cv2, yacs, ... and real Paddle features), so the packages on the way
to the hot-path modules are registered as empty namespace stubs and only these files execute:
    src/models/EMRT_utils/{utils,initializer,position_encoding,layers,transformer_encoder_decoder}.py
    src/models/backbones/swin_transformer.py      (Identity / DropPath / Mlp imported by t_e_d.py:17)
    src/models/paddle_EMRT.py                      (UpHead, paddle_EMRT.py:115-181)
    src/api/infer.py, src/utils/metrics.py
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("EMRT_REFERENCE_ROOT", "/root/reference/semantic_segmentation")
_PKG = "emrt_reference"          # private top-level name: never collides with our own packages
_loaded = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src/models/EMRT_utils/transformer_encoder_decoder.py"))


def _stub_package(name, path=None):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    m.__package__ = name
    sys.modules[name] = m
    return m


def _load(modname, relpath):
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """-> namespace with the reference's modules: ted (transformer_encoder_decoder), utils, layers, position_encoding,
    infer, metrics, emrt (paddle_EMRT)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_ROOT}")
    if os.environ.get("EMRT_USE_REAL_PADDLE") == "1":
        import paddle  # noqa: F401  (a real PaddlePaddle 2.1-2.4: INTEGRATION.md section 6)
    else:
        from . import paddle_on_torch
        paddle_on_torch.install()
    src = os.path.join(REF_ROOT, "src")
    # namespace stubs instead of the reference's import-everything __init__ files
    for name, path in ((_PKG, src), (f"{_PKG}.models", f"{src}/models"), (f"{_PKG}.models.EMRT_utils", f"{src}/models/EMRT_utils"),
                       (f"{_PKG}.models.backbones", f"{src}/models/backbones"), (f"{_PKG}.api", f"{src}/api"),
                       (f"{_PKG}.utils", f"{src}/utils")):
        _stub_package(name, path)
    # `from src.utils import load_pretrained_model` (swin_transformer.py:23) and the absolute imports of paddle_EMRT.py
    src_pkg = _stub_package("src", None)
    src_utils = _stub_package("src.utils", None)
    src_utils.load_pretrained_model = lambda *a, **kw: None
    src_pkg.utils = src_utils
    src_models = _stub_package("src.models", None)
    src_bb = _stub_package("src.models.backbones", None)
    src_bb.segformer_paddleSeg = types.ModuleType("segformer_paddleSeg")
    src_dec = _stub_package("src.models.decoders", None)
    fcn = _stub_package("src.models.decoders.fcn_head", None)
    fcn.FCNHead = type("FCNHead", (), {})
    src_pkg.models, src_models.backbones, src_models.decoders, src_dec.fcn_head = src_models, src_bb, src_dec, fcn
    bb = sys.modules[f"{_PKG}.models.backbones"]
    for missing in ("get_segmentation_backbone", "paddle_vision_resnet", "resnext", "resnest"):
        setattr(bb, missing, None)                   # names paddle_EMRT.py:5-8 imports; never called by UpHead
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")
    E = f"{_PKG}.models.EMRT_utils"
    _load(f"{_PKG}.models.backbones.swin_transformer", "src/models/backbones/swin_transformer.py")
    _loaded["initializer"] = _load(f"{E}.initializer", "src/models/EMRT_utils/initializer.py")
    _loaded["utils"] = _load(f"{E}.utils", "src/models/EMRT_utils/utils.py")
    _loaded["position_encoding"] = _load(f"{E}.position_encoding", "src/models/EMRT_utils/position_encoding.py")
    _loaded["layers"] = _load(f"{E}.layers", "src/models/EMRT_utils/layers.py")
    _loaded["ted"] = _load(f"{E}.transformer_encoder_decoder", "src/models/EMRT_utils/transformer_encoder_decoder.py")
    _loaded["infer"] = _load(f"{_PKG}.api.infer", "src/api/infer.py")
    _loaded["metrics"] = _load(f"{_PKG}.utils.metrics", "src/utils/metrics.py")
    _loaded["emrt"] = _load(f"{_PKG}.models.paddle_EMRT", "src/models/paddle_EMRT.py")
    return types.SimpleNamespace(**_loaded)


def load_params(layer, params, prefix=""):
    """Copies a {paddle state-dict key: ndarray} dict (oracle.make_*_params) into a reference Layer, in place."""
    import numpy as np
    import torch
    own = dict(layer.named_parameters())
    own.update(dict(layer.named_buffers()))
    missing = [k for k in own if prefix + k not in params]
    if missing:
        raise KeyError(f"no value for reference parameters {missing[:6]}")
    with torch.no_grad():
        for k, p in own.items():
            p.as_subclass(torch.Tensor).copy_(torch.as_tensor(np.asarray(params[prefix + k])).reshape(p.shape))
    return layer.eval()
