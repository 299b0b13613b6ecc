"""Import shim for the UNMODIFIED reference (simplify23/MRN) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py and the `-m "not gpu"` pinning tests to
check the oracle restatement against the reference's own modules.  /root/reference does not exist
on the GPU box; nothing under mrn_b200/, bench.py's product arm, or the `-m gpu` tests imports this.

Shims:
  * timm.models.layers.trunc_normal_  -> torch.nn.init.trunc_normal_   (modules/svtr.py:4)
  * lmdb, natsort, mmcv, nltk         -> import-time stubs             (data/dataset.py:6-8, test.py:13-14)
  * torch.Tensor.cuda                 -> identity on a CUDA-less host  (modules/svtr.py:119,125)
"""
import os
import sys
import types
import contextlib

REFERENCE_ROOT = os.environ.get("MRN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "model.py"))


def _install_stubs():
    import torch
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["timm.models.layers"] = layers
    for name in ("lmdb", "natsort", "mmcv", "nltk", "nltk.metrics", "nltk.metrics.distance"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["natsort"], "natsorted"):
        sys.modules["natsort"].natsorted = sorted
    if not hasattr(sys.modules["mmcv"], "Config"):
        sys.modules["mmcv"].Config = object
    dist = sys.modules["nltk.metrics.distance"]
    if not hasattr(dist, "edit_distance"):
        def edit_distance(a, b):
            prev = list(range(len(b) + 1))
            for i, ca in enumerate(a, 1):
                cur = [i]
                for j, cb in enumerate(b, 1):
                    cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
                prev = cur
            return prev[-1]
        dist.edit_distance = edit_distance
        sys.modules["nltk"].metrics = sys.modules["nltk.metrics"]
        sys.modules["nltk.metrics"].distance = dist


@contextlib.contextmanager
def reference_modules():
    """Context manager: puts the reference on sys.path and yields a namespace with its modules."""
    import torch
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    saved_cuda = torch.Tensor.cuda
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REFERENCE_ROOT)
    # the reference's top-level package names (modules, tools, data, il_modules) are generic: evict clashes
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k.split(".")[0] in ("modules", "tools", "data", "il_modules", "test")}
    try:
        import modules.model as ref_model
        import modules.dm_router as ref_router
        import modules.svtr as ref_svtr
        import tools.utils as ref_utils
        ns = types.SimpleNamespace(model=ref_model, dm_router=ref_router, svtr=ref_svtr, utils=ref_utils)
        yield ns
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in list(sys.modules):
            if k.split(".")[0] in ("modules", "tools", "data", "il_modules", "test"):
                sys.modules.pop(k)
        sys.modules.update(saved)
        torch.Tensor.cuda = saved_cuda
