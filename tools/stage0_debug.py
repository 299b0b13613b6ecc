"""Per-parameter gradient error of the stage-0 CUDA backward vs the reference golden (debug aid)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from test_gpu_stage0 import _run_step
from test_oracle_pinning import pview

PREC = int(os.environ.get("PREC", "0"))
for name in sys.argv[1:] or ["svtr_stage0_i2_b3"]:
    g, cc, sd, tp, logits, c, bn_train, pre = _run_step(name, PREC)
    print(name, "prec", PREC, "loss", float(c["loss"]), "ref", float(g["loss"]))
    tn = float(g["grad_total_norm"])
    for key, gg in tp.state(tp.grads).items():
        ref = g["grad." + pre + key]
        scale = max(float(np.abs(ref).max()), 1e-4 * tn)
        err = np.abs(pview(gg.contiguous().cpu(), g) - ref).max() / scale
        if err > (1e-4 if PREC == 0 else 3e-2) or os.environ.get('ALL'):
            print("%-70s err %.3e  |ref| %.3e |got| %.3e" % (key, err, float(g["gradnorm." + pre + key]), float(gg.norm())))
