"""Load the UNMODIFIED reference functions from /root/reference (build container only).

TEST INFRASTRUCTURE.  `tools/eval_mm_protocol.py` cannot be imported as-is: its
`from datasets.dataset import MultiModalDataset` (eval_mm_protocol.py:29) resolves to the
HuggingFace `datasets` package and `models.model` (eval_mm_protocol.py:30) pulls CLIP weights.
Neither is needed for the feature-level math, so both are stubbed before exec.
"""
import importlib.util
import io
import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("REID_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tools", "eval_mm_protocol.py"))


_cache = {}


def load_reference_eval():
    """Returns the reference module object for tools/eval_mm_protocol.py."""
    if "eval" in _cache:
        return _cache["eval"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    saved = {k: sys.modules.get(k) for k in ("datasets", "datasets.dataset", "models.model", "configs", "configs.config")}
    try:
        ds = types.ModuleType("datasets"); ds.__path__ = []
        dsd = types.ModuleType("datasets.dataset")
        dsd.MultiModalDataset = type("MultiModalDataset", (), {})
        mm = types.ModuleType("models.model")
        mm.CLIPBasedMultiModalReIDModel = type("CLIPBasedMultiModalReIDModel", (), {})
        cc = types.ModuleType("configs"); cc.__path__ = []
        ccc = types.ModuleType("configs.config")
        ccc.TrainingConfig = type("TrainingConfig", (), {})
        sys.modules.update({"datasets": ds, "datasets.dataset": dsd, "models.model": mm,
                            "configs": cc, "configs.config": ccc})
        spec = importlib.util.spec_from_file_location(
            "ref_eval_mm_protocol", os.path.join(REFERENCE_ROOT, "tools", "eval_mm_protocol.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    # silence tqdm progress bars of the reference loop
    mod.tqdm = lambda it, **kw: it
    _cache["eval"] = mod
    return mod


def load_reference_sdm():
    """Returns the reference `sdm_loss_stable` (models/sdm_loss.py:13), stdout silenced by caller."""
    if "sdm" in _cache:
        return _cache["sdm"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    spec = importlib.util.spec_from_file_location(
        "ref_sdm_loss", os.path.join(REFERENCE_ROOT, "models", "sdm_loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache["sdm"] = mod
    return mod


def quiet(fn, *a, **kw):
    """Call fn with stdout swallowed (the reference prints warnings / debug lines)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def load_reference_train_eval():
    """The UNMODIFIED `compute_map`, `compute_cmc`, `_reid_map`, `_flatten_loaders`, `evaluate_one_query` and
    `validate_competition_style` of train.py (:101-138, :402-424, :451-631).  train.py cannot be imported (dataset /
    model / CLIP imports at module level), so the function definitions are cut out of its source with `ast` and
    executed as they stand in a namespace that holds what they use (torch, F, np, os, hashlib, pickle, DataLoader,
    Subset); `_extract_feats_and_ids` -- the model forward -- is the only stub."""
    if "train_eval" in _cache:
        return _cache["train_eval"]
    import ast
    import numpy as np
    import torch
    import torch.nn.functional as F
    path = os.path.join(REFERENCE_ROOT, "train.py")
    if not os.path.isfile(path):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    import hashlib
    import pickle
    from torch.utils.data import DataLoader, Subset

    def _extract_feats_and_ids(model, loader, device):
        """STUB of train.py:428-449 (the one step that needs the model): serves the pre-extracted (feature, id) items the
        loader's dataset holds, for the unmodified evaluate_one_query / validate_competition_style below."""
        ds = loader.dataset
        items = [ds[i] for i in range(len(ds))]
        return torch.stack([f for f, _ in items]), torch.stack([torch.as_tensor(i) for _, i in items])

    ns = {"torch": torch, "F": F, "np": np, "os": os, "hashlib": hashlib, "pickle": pickle, "DataLoader": DataLoader,
          "Subset": Subset, "_extract_feats_and_ids": _extract_feats_and_ids}
    wanted = ("compute_map", "compute_cmc", "_reid_map", "_flatten_loaders", "evaluate_one_query", "validate_competition_style")
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    _cache["train_eval"] = ns
    return ns


def load_reference_compute_loss():
    """The UNMODIFIED `CLIPBasedMultiModalReIDModel.compute_loss` (models/model.py:512-659) as a plain function
    `compute_loss(self, outputs, labels)`.  models/model.py cannot be imported (its CLIP backbone needs the network), so
    the method definition is cut out of the class with `ast` and executed as it stands; its function-local
    `from .sdm_loss import sdm_loss_stable` (:556) resolves to the reference's own models/sdm_loss.py.  `self` only needs
    the attributes the method reads: ce_loss, current_epoch, config.sdm_weight_warmup_epochs, contrastive_weight,
    ce_weight, sdm_temperature, training."""
    if "compute_loss" in _cache:
        return _cache["compute_loss"]
    import ast
    import logging
    import typing
    import torch
    path = os.path.join(REFERENCE_ROOT, "models", "model.py")
    if not os.path.isfile(path):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    tree = ast.parse(open(path, encoding="utf-8").read())
    node = None
    for cls in tree.body:
        if isinstance(cls, ast.ClassDef) and cls.name == "CLIPBasedMultiModalReIDModel":
            for item in cls.body:
                if isinstance(item, ast.FunctionDef) and item.name == "compute_loss":
                    node = item
    if node is None:
        raise RuntimeError("compute_loss not found in %s" % path)
    sdm_mod = load_reference_sdm()
    pkg = types.ModuleType("refmodels"); pkg.__path__ = []
    ns = {"torch": torch, "logger": logging.getLogger("reference.models.model"), "Dict": typing.Dict, "List": typing.List,
          "Any": typing.Any, "Optional": typing.Optional, "__name__": "refmodels.model", "__package__": "refmodels"}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    fn = ns["compute_loss"]

    def call(self, outputs, labels):
        saved = {k: sys.modules.get(k) for k in ("refmodels", "refmodels.sdm_loss")}
        sys.modules["refmodels"] = pkg
        sys.modules["refmodels.sdm_loss"] = sdm_mod
        try:
            return quiet(fn, self, outputs, labels)
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    _cache["compute_loss"] = call
    return call
