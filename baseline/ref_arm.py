#!/usr/bin/env python
"""Comparison arm: the UNMODIFIED reference (baseline/_ref/, staged by baseline/install_ref.py) running its own
stage-1 loop -- il_modules/mrn.py::MRN._update_representation (:298-384) over modules/model.py::MRNNet.cross_forward
(:397-423), torch.nn.CTCLoss, CrossEntropyLoss, clip_grad_norm_, Adam, OneCycleLR -- on the bench workload
(SVTR-MRN, 6 experts, MLT17 class counts, B samples of synthetic 32x256x4 crops per iteration).

    python baseline/ref_arm.py --device cpu  --batch 256 --steps 3 --warmup 1      # CPU arm (bench.py --impl reference)
    python baseline/ref_arm.py --device cuda --batch 256 --steps 10 --warmup 3     # eager torch on the B200

Nothing of mrn_b200's kernels, models or engine is on this path: the reference's learner is built through its own
constructors (build_model / change_model x5, random init) and driven through its own training loop.  The only things
supplied from outside are what tiny_train.py would supply: the option namespace (config/svtr_mrn.py), a train loader
object with get_batch2(), a (tiny) validation loader, and import stubs for packages absent from this image
(timm.trunc_normal_, lmdb, natsort, mmcv, nltk -- none of them is on the stage-1 compute path).

Timing: the loader's get_batch2() is the iteration boundary; iteration k's time = t(call k+1) - t(call k) (with a
device synchronize on CUDA).  Iteration 1 and the last iteration (which run the reference's val()) lie outside the
timed window.  Prints ONE JSON line.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
CLASS_COUNTS = (1899, 2224, 3844, 4968, 5041, 5153)


def _stubs(torch):
    if "timm" not in sys.modules:
        try:
            import timm  # noqa: F401
        except Exception:
            timm = types.ModuleType("timm"); models = types.ModuleType("timm.models"); layers = types.ModuleType("timm.models.layers")
            layers.trunc_normal_ = torch.nn.init.trunc_normal_
            timm.models = models; models.layers = layers
            sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
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


def _load_config(path):
    ns = {}
    exec(compile(open(path).read(), path, "exec"), ns)
    opt = {}
    for k in ("common", "model", "optimizer", "train"):        # tiny_train.py:413-422 merges the four dicts
        opt.update(ns[k])
    return argparse.Namespace(**opt)


def synth_batch(B, seed, charset):
    """Same rule as mrn_b200/synth.py::synth_batch, restated here so this file has no dependency on the
    product package: images randn.clamp(-1,1); lengths 1..25; label ids in [2, C); domain ids in [0, I)."""
    import numpy as np
    import torch
    import zlib
    C = CLASS_COUNTS[-1]

    def rng(name):
        return np.random.default_rng([seed, zlib.crc32(name.encode())])
    img = torch.from_numpy(rng("image").standard_normal((B, 4, 32, 256), dtype=np.float32)).clamp_(-1, 1)
    g = rng("labels")
    lens = g.integers(1, 26, size=B)
    words = []
    for b in range(B):
        ids = g.integers(2, C, size=lens[b])
        # index 2 = [UNK] (any character outside the dict), 3 = ' ', >= 4 = charset[id - 4]  (tools/utils.py:15-31)
        words.append("".join("" if i == 2 else (" " if i == 3 else charset[i - 4]) for i in ids))
    dom = g.integers(0, len(CLASS_COUNTS), size=B)
    return img, words, [int(v) for v in dom]


def run(device, B, steps, warmup, threads, seed=111):
    if device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""             # the reference picks "cuda" whenever one is visible
    import torch
    if not os.path.isfile(os.path.join(REF, "modules", "model.py")):
        return {"unavailable": "baseline/_ref is not staged (run python baseline/install_ref.py where /root/reference exists)"}
    if device == "cuda" and not torch.cuda.is_available():
        return {"unavailable": "no CUDA device for the eager reference"}
    if threads:
        torch.set_num_threads(threads)
    _stubs(torch)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self       # modules/svtr.py:119,125 call .cuda() in the constructor
    sys.path.insert(0, REF)
    for k in list(sys.modules):
        if k.split(".")[0] in ("modules", "tools", "data", "il_modules", "test"):
            sys.modules.pop(k)
    import torch.backends.cudnn as cudnn
    opt = _load_config(os.path.join(REF, "config", "svtr_mrn.py"))
    n_total = warmup + steps + 2                 # iteration 1 (+val) | warm-up | timed | last iteration (+val)
    opt.num_iter = 2 * n_total                   # _update_representation runs num_iter // 2 iterations
    opt.val_interval = 10 ** 9
    opt.exp_name = "ref_arm"
    opt.batch_size = B
    # tiny_train.py:425-432
    import random
    import numpy as np
    random.seed(opt.manual_seed); np.random.seed(opt.manual_seed)
    torch.manual_seed(opt.manual_seed); torch.cuda.manual_seed_all(opt.manual_seed)
    cudnn.benchmark = True
    cudnn.deterministic = True

    work = tempfile.mkdtemp(prefix="ref_arm_")
    cwd = os.getcwd()
    os.chdir(work)
    os.makedirs(os.path.join("saved_models", opt.exp_name), exist_ok=True)
    try:
        from il_modules.mrn import MRN
        charset = [chr(0x4E00 + i) for i in range(CLASS_COUNTS[-1] - 4)]
        learner = MRN(opt)
        for i, c in enumerate(CLASS_COUNTS):                 # what incremental_train does per task (il_modules/mrn.py:136-157)
            learner._total_classes = c
            if i == 0:
                learner.build_model()
            else:
                learner.change_model()
        learner.character = charset
        learner.converter = learner.build_converter()
        assert learner._total_classes == CLASS_COUNTS[-1]
        for p in learner.model.module.model.parameters():    # experts 0..I-2 frozen by incremental_train, I-1 by update_step1
            p.requires_grad = False
        taski_log = []
        ce = torch.nn.CrossEntropyLoss(reduction="mean").to(learner.device)

        def taski_criterion(gate, idx):                      # il_modules/mrn.py:150-152, recording its value
            v = ce(gate, idx)
            taski_log.append(v.detach())
            return v
        learner.taski_criterion = taski_criterion

        batches = [synth_batch(B, 1000 + k, charset) for k in range(4)]
        if device == "cuda":
            batches = [(b[0].pin_memory(), b[1], b[2]) for b in batches]

        class Loader:
            def __init__(self):
                self.t = []

            def get_batch2(self):
                if device == "cuda":
                    torch.cuda.synchronize()
                self.t.append(time.perf_counter())
                return batches[(len(self.t) - 1) % len(batches)]
        loader = Loader()
        vimg, vwords, _ = synth_batch(2, 77, charset)
        valid = [(vimg, vwords)]
        learner._update_representation(0, len(CLASS_COUNTS) - 1, loader, valid)
        if device == "cuda":
            torch.cuda.synchronize()
        t = loader.t
        assert len(t) == n_total
        # iteration k (1-based) spans t[k-1] .. t[k]; timed iterations: 2+warmup .. 1+warmup+steps
        a, b = 1 + warmup, 1 + warmup + steps
        dt = t[b] - t[a]
        log = open(os.path.join("saved_models", opt.exp_name, "log_train.txt")).read()
        loss_first = None
        for line in log.splitlines():
            if line.startswith("[1/") and "Train_loss_clf" in line:
                loss_first = float(line.split("Train_loss_clf:")[1].split(",")[0])
                break
        out = {
            "device": device, "batch": B, "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1000.0,
            "samples_per_s": B * steps / dt, "threads": torch.get_num_threads(), "torch": torch.__version__,
            "loss_clf_first": loss_first, "taski_loss_first": float(taski_log[0]) if taski_log else None,
            "taski_loss_last": float(taski_log[-1]) if taski_log else None,
            "code": "il_modules/mrn.py::MRN._update_representation over modules/model.py::MRNNet.cross_forward (baseline/_ref, unmodified)",
            "flags": {"dtype": "fp32", "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32),
                      "cudnn_allow_tf32": bool(cudnn.allow_tf32), "cudnn_benchmark": True, "cudnn_deterministic": True,
                      "data_parallel_devices": torch.cuda.device_count() if device == "cuda" else 0},
        }
        if device == "cuda":
            out["gpu_name"] = torch.cuda.get_device_name(0)
            out["peak_mem_gb"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
        return out
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    real_stdout = sys.stdout
    sys.stdout = sys.stderr                        # the reference prints progress; keep stdout for the JSON line
    try:
        res = run(a.device, a.batch, a.steps, a.warmup, a.threads or (os.cpu_count() or 1))
    except Exception as ex:                        # noqa: BLE001
        import traceback
        traceback.print_exc()
        res = {"unavailable": "reference arm failed: %s: %s" % (type(ex).__name__, ex)}
    sys.stdout = real_stdout
    print(json.dumps(res))
