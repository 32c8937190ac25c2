"""Loader for the UNMODIFIED reference nets (test infrastructure, container-only).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``oracle/make_golden.py`` and
``tests/test_oracle_vs_reference.py`` import this module, and only in the build
container where ``/root/reference`` exists (it does not exist on the GPU box).

The reference package ``deeplio`` imports third-party modules that are absent
here (``open3d``, ``matplotlib``, ``liegroups``, ``tensorboardX``,
``pytorch_model_summary``: deeplio/common/__init__.py:1-4,
deeplio/models/misc.py:5, deeplio/models/worker.py).  None of them is used by
the four nn.Module subsystems under deeplio/models/nets, so empty stub modules
are inserted into ``sys.modules`` before the import.  No reference source is
copied; the modules are executed where they lie.
"""
import argparse
import contextlib
import copy
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("DEEPLIO_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "open3d", "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "liegroups",
    "liegroups.torch", "tensorboardX", "pytorch_model_summary", "PIL", "PIL.Image",
]


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deeplio", "models", "nets"))


def _install_stubs():
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            if name in ("PIL", "PIL.Image"):
                __import__(name)
                continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__dict__.setdefault("__path__", [])
        sys.modules[name] = mod
    # attributes looked up at import time (deeplio/models/misc.py:5 etc.)
    lt = sys.modules["liegroups.torch"]
    for attr in ("SO3", "SE3", "utils"):
        if not hasattr(lt, attr):
            setattr(lt, attr, types.SimpleNamespace())
    sys.modules["liegroups"].torch = lt
    tb = sys.modules["tensorboardX"]
    if not hasattr(tb, "SummaryWriter"):
        tb.SummaryWriter = object
    pms = sys.modules["pytorch_model_summary"]
    if not hasattr(pms, "summary"):
        pms.summary = lambda *a, **k: ""
    mpl = sys.modules["matplotlib"]
    mpl.cm = sys.modules["matplotlib.cm"]
    mpl.pyplot = sys.modules["matplotlib.pyplot"]


@contextlib.contextmanager
def _in_tmp_cwd():
    """The reference logger opens ``deeplio.txt`` in the cwd (common/logger.py:69)."""
    old = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="dlio_ref_")
    os.chdir(tmp)
    try:
        yield
    finally:
        os.chdir(old)


def import_reference():
    """Import ``deeplio.models.nets`` and ``deeplio.models.misc`` from the reference tree."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with _in_tmp_cwd():
        import deeplio.models.nets as nets  # noqa
        import deeplio.models.misc as misc  # noqa
    return nets, misc


def patch_cfg(cfg):
    """Oracle patch P1: the shipped config.yaml lacks
    ``deeplio.fusion-net.pretrained`` (KeyError at nets/__init__.py:181)."""
    cfg = copy.deepcopy(cfg)
    fn = cfg["deeplio"].get("fusion-net")
    if fn is not None:
        fn.setdefault("pretrained", False)
    return cfg


def build_reference_model(cfg, height, width, batch_size=1, device="cpu"):
    """``nets.get_model`` exactly as trainer.py:52 calls it (after worker.py:41)."""
    nets, misc = import_reference()
    cfg = patch_cfg(cfg)
    args = argparse.Namespace(device=device, batch_size=batch_size)
    misc.build_config_container(cfg, args)
    with _in_tmp_cwd():
        model = nets.get_model(input_shape=(3, height, width), cfg=cfg, device=device)
    return model


def apply_patch_p4(model):
    """Oracle patch P4: DeepLIOFusionSoft multiplies its inputs in place
    (fusion_nets.py:72-73), which makes autograd raise for lidar nets whose last op is a ReLU with
    no dropout after it.  Same arithmetic, out of place, bound on the instance only."""
    import torch

    fn = getattr(model, "fusion_net", None)
    if fn is None or not hasattr(fn, "layers"):
        return model

    def forward(x, _fn=fn):
        lidar_feat, imu_feat = x[0], x[1]
        cat_feat = torch.cat((lidar_feat, imu_feat), dim=2)
        _fn.s1_feat = torch.sigmoid(_fn.layers[0](cat_feat))
        _fn.s2_feat = torch.sigmoid(_fn.layers[1](cat_feat))
        return torch.cat((lidar_feat * _fn.s1_feat, imu_feat * _fn.s2_feat), dim=2)

    fn.forward = forward
    return model
