#!/usr/bin/env python
"""The bench.py headline step (SVTR-MRN, 6 experts, B = 256, bf16, train-mode experts) between cudaProfilerStart / Stop,
eager launches, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py [--steps 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import CLASS_COUNTS, make_opt  # noqa: E402
from mrn_b200 import synth  # noqa: E402
from mrn_b200.il_modules.mrn import MRN, RankLocal, FusedAdam  # noqa: E402
from mrn_b200.modules.model import MRNNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--mode", default="train", choices=["train", "infer"])
a = ap.parse_args()
opt = make_opt("bf16", 0, "svtr")
net = MRNNet(opt)
for c in CLASS_COUNTS:
    net.update_fc(opt.hidden_size, c)
    net.build_prediction(opt, c)
net.load_state_dict(synth.ctor_state_dict(CLASS_COUNTS, 111), strict=True)
net = net.cuda()
learner = MRN(opt)
learner.model = RankLocal(net)
learner.model.train() if a.mode == "train" else learner.model.eval()
learner.optimizer = FusedAdam(net, opt.lr, 20000, grad_clip=5, schedule="super")
batches = [tuple(t.cuda() for t in synth.synth_batch(a.batch, CLASS_COUNTS, 1000 + k)) for k in range(2)]


def step(k):
    img, tgt, lens, dom = batches[k % 2]
    if a.mode == "train":
        return learner.train_step_stage1(img, tgt, lens, dom)
    return learner.infer_batch(img, "TF")


for k in range(3):
    step(k)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for k in range(a.steps):
    out = step(k)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profile_step done")
