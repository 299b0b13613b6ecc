import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from oracle import synth
from mrn_b200 import ops
cc, B, seed = (53,), 64, 9
sd = synth.synth_state_dict(cc, seed, arch="crnn")
img, tgt, lens, dom = synth.synth_batch(B, cc, seed)
esd = {k[len("model.0."):]: v for k, v in sd.items() if k.startswith("model.0.")}
x, t, l = img.cuda(), tgt.cuda(), lens.cuda()
res = []
for prec in (0, 1):
    tp = ops.CrnnTrainPack(esd, "cuda", prec)
    logits = ops.crnn_train_forward(tp, x, True, True)
    rr = ops.gate_combine([logits], torch.ones(B, 1, device="cuda"), t, l)
    c = ops.ctc_lattice(rr["lpe"], t, l, want_occ=True)
    dlogits = ops.ctc_dense_grad(logits, rr["lse"], c["occ"], c["nll"], t, l, 1.0 / B)
    ops.crnn_train_backward(tp, dlogits, B, True)
    torch.cuda.synchronize()
    res.append((float(c["loss"]), logits.clone(), {k: v.clone() for k, v in tp.state(tp.grads).items()}))
(l32, z32, g32), (l16, z16, g16) = res
print("loss", l32, l16, "logits rel", float((z16 - z32).abs().max() / z32.abs().max()))
for k in g32:
    e = float((g16[k].double() - g32[k].double()).norm()) / max(float(g32[k].double().norm()), 1e-30)
    print("%-60s %.4f  |g| %.3e" % (k, e, float(g32[k].norm())))
