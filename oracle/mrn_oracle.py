"""CPU oracle: a plain restatement of the MRN multiplexed-routing train / infer step.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path (mrn_b200/) never does, and has no CPU fallback.

Parity status: the reference (simplify23/MRN) ships NO tests, golden vectors or known-answer files for
this path, so the oracle is pinned against *outputs of the reference itself run in
the build container*: oracle/make_golden.py imports the unmodified reference modules from
/root/reference, loads a deterministic synthetic state_dict (oracle/synth.py), and stores the reference's
outputs under tests/golden/.  tests/test_oracle_pinning.py checks this restatement against those
fixtures (always) and against the live reference (when /root/reference is present).  The arithmetic lives
in a third-party dependency, PyTorch (reference README pins torch 1.6.0 / 1.9.1+cu111; here 2.11.0);
the CTC recursion is additionally restated from scratch (ctc_nll_and_grad) and checked against
torch.nn.functional.ctc_loss in fp64 and a brute-force path enumeration.

Everything here is functional: weights come from a state_dict with the reference's key names
(`model.{i}.model.FeatureExtraction.ConvNet.*`, `model.{i}.model.SequenceModeling.0.*`, `model.{i}.fc.*`,
`route.*`, `channel_route.*`, `dm_router.0.*`), tensors are torch CPU tensors (fp32 or fp64).
Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# SVTR expert  (modules/svtr.py, modules/model.py:82-101,133-148)
# ----------------------------------------------------------------------------------------------

SVTR_DIMS = (64, 128, 256)          # modules/svtr.py:319 embed_dim
SVTR_DEPTH = (3, 6, 3)              # modules/svtr.py:320 depth
SVTR_HEADS = (2, 4, 8)              # modules/svtr.py:321 num_heads
SVTR_MIXER = ("Local",) * 6 + ("Global",) * 6   # modules/svtr.py:322-323
SVTR_LOCAL_K = (7, 11)              # modules/svtr.py:324
SVTR_GRID = ((8, 64), (4, 64), (2, 64))         # HW per stage for a 32x256 input (modules/svtr.py:348,398,425)
SVTR_DROP_PATH_RATE = 0.1           # modules/svtr.py:332


def svtr_drop_path_rates() -> List[float]:
    """modules/svtr.py:382  dpr = np.linspace(0, drop_path_rate, sum(depth))."""
    n = sum(SVTR_DEPTH)
    return [SVTR_DROP_PATH_RATE * i / (n - 1) for i in range(n)]


def local_mask(H: int, W: int, hk: int = 7, wk: int = 11, dtype=torch.float32) -> torch.Tensor:
    """Additive 0/-inf mask of the Local mixer (modules/svtr.py:116-128).

    allowed(h,w ; h',w') = |h-h'| <= hk//2 and |w-w'| <= wk//2  (SURVEY.md Appendix A.1)."""
    hh = torch.arange(H).view(H, 1, 1, 1)
    ww = torch.arange(W).view(1, W, 1, 1)
    h2 = torch.arange(H).view(1, 1, H, 1)
    w2 = torch.arange(W).view(1, 1, 1, W)
    ok = ((hh - h2).abs() <= hk // 2) & ((ww - w2).abs() <= wk // 2)
    m = torch.full((H * W, H * W), float("-inf"), dtype=dtype)
    m[ok.reshape(H * W, H * W)] = 0.0
    return m


def _bn(x, sd, p, mode: str, eps=1e-5):
    """nn.BatchNorm2d (modules/svtr.py:229,232). mode 'eval' = running stats, 'batch' = batch stats
    (what a module in .train() computes; reference quirk 4, il_modules/mrn.py:401)."""
    w, b = sd[p + "weight"], sd[p + "bias"]
    if mode == "eval":
        mean, var = sd[p + "running_mean"], sd[p + "running_var"]
    else:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
    xh = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps)
    return xh * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def _gelu(x):
    """nn.GELU() exact erf form (modules/svtr.py:230; modules/dm_router.py:42)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _ln(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def svtr_attention(x, sd, p, heads: int, mask: Optional[torch.Tensor]):
    """Attention.forward (modules/svtr.py:133-152): qkv bias on, q scaled by head_dim^-0.5,
    additive local mask, softmax over keys, proj."""
    Bn, N, C = x.shape
    hd = C // heads
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    qkv = qkv.reshape(Bn, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-1, -2)
    if mask is not None:
        attn = attn + mask.to(attn.dtype)
    attn = torch.softmax(attn, dim=-1)
    o = (attn @ v).permute(0, 2, 1, 3).reshape(Bn, N, C)
    return F.linear(o, sd[p + "proj.weight"], sd[p + "proj.bias"])


def svtr_block(x, sd, p, heads, mask, drop_scale: Optional[torch.Tensor]):
    """Block.forward (modules/svtr.py:200-204), pre-norm, LN eps 1e-6 (:348), DropPath per sample.

    drop_scale: None (eval / rate 0) or [2,B] multipliers (0 or 1/keep_prob) for the mixer and MLP
    branches (modules/svtr.py:7-22)."""
    h = svtr_attention(_ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6), sd, p + "mixer.", heads, mask)
    if drop_scale is not None:
        h = h * drop_scale[0].view(-1, 1, 1).to(h.dtype)
    x = x + h
    h = _ln(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = F.linear(_gelu(h), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    if drop_scale is not None:
        h = h * drop_scale[1].view(-1, 1, 1).to(h.dtype)
    return x + h


def svtr_subsample(x, sd, p, C, H, W):
    """SubSample 'Conv' (modules/svtr.py:277,306-308): conv3x3 stride (2,1) pad 1 -> tokens -> LN(eps 1e-5)."""
    B = x.shape[0]
    x = x.transpose(1, 2).reshape(B, C, H, W)
    x = F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=(2, 1), padding=1)
    x = x.flatten(2).transpose(1, 2)
    return _ln(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)


def svtr_backbone(sd: Dict[str, torch.Tensor], p: str, image: torch.Tensor, bn_mode: str = "eval",
                  drop_scales: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SVTR.forward_features (modules/svtr.py:500-528) followed by Model_Extractor's permute / avg-pool /
    squeeze (modules/model.py:88-95; the pooled H axis has size 1) -> visual feature [B, 64, 512].

    drop_scales: None or [12, 2, B]."""
    B = image.shape[0]
    x = F.conv2d(image, sd[p + "patch_embed.proj.0.weight"], sd[p + "patch_embed.proj.0.bias"], stride=2, padding=1)
    x = _gelu(_bn(x, sd, p + "patch_embed.proj.1.", bn_mode))
    x = F.conv2d(x, sd[p + "patch_embed.proj.3.weight"], sd[p + "patch_embed.proj.3.bias"], stride=2, padding=1)
    x = _gelu(_bn(x, sd, p + "patch_embed.proj.4.", bn_mode))
    x = x.flatten(2).transpose(1, 2) + sd[p + "pos_embed"]                      # svtr.py:246-254,511
    blk = 0
    for stage in range(3):
        H, W = SVTR_GRID[stage]
        for j in range(SVTR_DEPTH[stage]):
            mask = local_mask(H, W, *SVTR_LOCAL_K, dtype=x.dtype) if SVTR_MIXER[blk] == "Local" else None
            ds = None if drop_scales is None else drop_scales[blk]
            x = svtr_block(x, sd, f"{p}blocks{stage + 1}.{j}.", SVTR_HEADS[stage], mask, ds)
            blk += 1
        x = svtr_subsample(x, sd, f"{p}sub_sample{stage + 1}.", SVTR_DIMS[stage], H, W)
    return x        # [B, 64, 512]  (svtr.py:527 + model.py:88-95 are pure relabelings of this tensor)


# ----------------------------------------------------------------------------------------------
# CRNN expert: VGG feature extractor + 2 x BidirectionalLSTM
# (modules/feature_extraction.py:19-47, modules/sequence_modeling.py:4-22, modules/model.py:46-57,82-101)
# ----------------------------------------------------------------------------------------------

def vgg_backbone(sd: Dict[str, torch.Tensor], p: str, image: torch.Tensor, bn_mode: str = "eval") -> torch.Tensor:
    """VGG_FeatureExtractor.forward (feature_extraction.py:19-47: Sequential indices 0..19) followed by
    Model_Extractor's permute / AdaptiveAvgPool2d((None,1)) / squeeze (model.py:88-95) -> [B, 63, 512]."""
    x = F.max_pool2d(F.relu(F.conv2d(image, sd[p + "0.weight"], sd[p + "0.bias"], padding=1)), 2, 2)      # 0-2
    x = F.max_pool2d(F.relu(F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], padding=1)), 2, 2)          # 3-5
    x = F.relu(F.conv2d(x, sd[p + "6.weight"], sd[p + "6.bias"], padding=1))                              # 6-7
    x = F.max_pool2d(F.relu(F.conv2d(x, sd[p + "8.weight"], sd[p + "8.bias"], padding=1)), (2, 1), (2, 1))   # 8-10
    x = F.relu(_bn(F.conv2d(x, sd[p + "11.weight"], None, padding=1), sd, p + "12.", bn_mode))            # 11-13
    x = F.relu(_bn(F.conv2d(x, sd[p + "14.weight"], None, padding=1), sd, p + "15.", bn_mode))            # 14-16
    x = F.max_pool2d(x, (2, 1), (2, 1))                                                                   # 17
    x = F.relu(F.conv2d(x, sd[p + "18.weight"], sd[p + "18.bias"]))                                       # 18-19  [B,512,1,63]
    return x.permute(0, 3, 1, 2).mean(dim=3)                                                              # model.py:88-95


def lstm_direction(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> torch.Tensor:
    """One direction of nn.LSTM(batch_first=True), zero initial state, gate order i, f, g, o:
    c_t = sigmoid(f) * c_{t-1} + sigmoid(i) * tanh(g);  h_t = sigmoid(o) * tanh(c_t).   x [B,T,K] -> [B,T,H]."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    pre = F.linear(x, w_ih, b_ih + b_hh)
    h = torch.zeros(B, H, dtype=x.dtype)
    c = torch.zeros(B, H, dtype=x.dtype)
    out = torch.empty(B, T, H, dtype=x.dtype)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        a = pre[:, t] + h @ w_hh.t()
        i_, f_, g_, o_ = a[:, :H], a[:, H:2 * H], a[:, 2 * H:3 * H], a[:, 3 * H:]
        c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
        h = torch.sigmoid(o_) * torch.tanh(c)
        out[:, t] = h
    return out


def bilstm(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """BidirectionalLSTM.forward (sequence_modeling.py:12-22): [fwd | bwd] hidden states -> Linear(2H -> out)."""
    f = lstm_direction(x, sd[p + "rnn.weight_ih_l0"], sd[p + "rnn.weight_hh_l0"], sd[p + "rnn.bias_ih_l0"],
                       sd[p + "rnn.bias_hh_l0"], False)
    b = lstm_direction(x, sd[p + "rnn.weight_ih_l0_reverse"], sd[p + "rnn.weight_hh_l0_reverse"],
                       sd[p + "rnn.bias_ih_l0_reverse"], sd[p + "rnn.bias_hh_l0_reverse"], True)
    return F.linear(torch.cat([f, b], dim=-1), sd[p + "linear.weight"], sd[p + "linear.bias"])


def expert_arch(sd, i: int = 0) -> str:
    return "crnn" if f"model.{i}.model.SequenceModeling.0.rnn.weight_ih_l0" in sd else "svtr"


def expert_forward(sd, i: int, image, bn_mode="eval", drop_scales=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Model.forward with the CTC head (modules/model.py:133-148,75-80,176-181): SVTR + Linear 'None' sequence
    model, or VGG + two BidirectionalLSTMs (picked from the state_dict keys).

    Returns (feature [B,T,256], predict [B,T,C_i])."""
    if expert_arch(sd, i) == "crnn":
        vis = vgg_backbone(sd, f"model.{i}.model.FeatureExtraction.ConvNet.", image, bn_mode)
        feat = bilstm(sd, f"model.{i}.model.SequenceModeling.1.", bilstm(sd, f"model.{i}.model.SequenceModeling.0.", vis))
    else:
        vis = svtr_backbone(sd, f"model.{i}.model.FeatureExtraction.ConvNet.", image, bn_mode, drop_scales)
        feat = F.linear(vis, sd[f"model.{i}.model.SequenceModeling.0.weight"], sd[f"model.{i}.model.SequenceModeling.0.bias"])
    pred = F.linear(feat, sd[f"model.{i}.fc.weight"], sd[f"model.{i}.fc.bias"])
    return feat, pred


# ----------------------------------------------------------------------------------------------
# DM-Router + gate head + combine  (modules/dm_router.py:50-67, modules/model.py:361-423)
# ----------------------------------------------------------------------------------------------

def dm_router(sd, x: torch.Tensor, p: str = "dm_router.0.") -> torch.Tensor:
    """DM_Router.forward (modules/dm_router.py:50-67) written out as in SURVEY.md Appendix A.2.
    x [B,I,T,D] -> [B,I,T,D]."""
    B, I, T, D = x.shape
    h = _gelu(F.linear(_ln(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5),
                       sd[p + "proj_1.weight"], sd[p + "proj_1.bias"]))              # :55-57
    h = h.reshape(B, I * T, 2 * D)                                                   # :58
    u, v = h[..., :D], h[..., D:]                                                    # :12
    v = _ln(v, sd[p + "spatial_gating.norm.weight"], sd[p + "spatial_gating.norm.bias"], 1e-5)   # :13
    v = torch.einsum("nm,bmc->bnc", sd[p + "spatial_gating.proj.weight"], v) \
        + sd[p + "spatial_gating.proj.bias"].view(1, -1, 1)                          # :14-16
    y = F.linear(u * v, sd[p + "proj_2.weight"], sd[p + "proj_2.bias"]).reshape(B, I, T, D) + x   # :17,60-62
    q = y.permute(0, 1, 3, 2).reshape(B, I * D, T)                                   # :63  b (d c) p
    g = _ln(q, sd[p + "channel_gating.norm.weight"], sd[p + "channel_gating.norm.bias"], 1e-5)   # :29
    g = torch.einsum("kj,bjt->bkt", sd[p + "channel_gating.proj.weight"], g) \
        + sd[p + "channel_gating.proj.bias"].view(1, -1, 1)                          # :30-32
    q = (q * g).reshape(B, I, D, T).permute(0, 1, 3, 2)                              # :33,65
    return F.linear(q, sd[p + "proj_3.weight"], sd[p + "proj_3.bias"]) + x           # :66-67


def gate_scores(sd, router_out: torch.Tensor) -> torch.Tensor:
    """modules/model.py:402-405 (train) / 371-375 (eval): rearrange 'b h w c -> b w (h c)' -> channel_route
    -> permute -> route -> squeeze.  Returns r [B,I]."""
    B, I, T, D = router_out.shape
    z = router_out.permute(0, 2, 1, 3).reshape(B, T, I * D)
    s = F.linear(z, sd["channel_route.weight"], sd["channel_route.bias"])            # [B,T,I]
    r = F.linear(s.permute(0, 2, 1), sd["route.weight"], sd["route.bias"])           # [B,I,1]
    return r.squeeze(-1)


def pad_ones(z: torch.Tensor, C: int) -> torch.Tensor:
    """MRNNet.pad_zeros_features (modules/model.py:361-364): pads with ONES, not zeros."""
    B, T, Ci = z.shape
    if Ci == C:
        return z
    return torch.cat([z, torch.ones(B, T, C - Ci, dtype=z.dtype)], dim=-1)


def combine_soft(preds: Sequence[torch.Tensor], gate: torch.Tensor) -> torch.Tensor:
    """modules/model.py:410-423: logits[b,t,c] = sum_i gate[b,i] * pad_i[b,t,c]."""
    C = preds[-1].shape[-1]
    out = torch.zeros_like(preds[-1])
    for i, z in enumerate(preds):
        out = out + gate[:, i].view(-1, 1, 1) * pad_ones(z, C)
    return out


def combine_hard(preds: Sequence[torch.Tensor], index: torch.Tensor) -> torch.Tensor:
    """modules/model.py:383-393: logits[b] = pad_{index[b]}[b]."""
    C = preds[-1].shape[-1]
    padded = [pad_ones(z, C) for z in preds]
    return torch.stack([padded[int(index[b])][b] for b in range(index.shape[0])], 0)


def mrn_forward(sd, n_experts: int, image, cross=True, is_train=True, bn_mode="eval", drop_scales=None):
    """MRNNet.forward (modules/model.py:343-359).  drop_scales: None or [I,12,2,B].

    Returns dict(logits, index, features [B,I,T,D], router_out, scores r[B,I], preds list)."""
    if not cross:                                                                    # :346-348
        ds = None if drop_scales is None else drop_scales[n_experts - 1]
        feat, pred = expert_forward(sd, n_experts - 1, image, bn_mode, ds)
        return dict(logits=pred, index=None, preds=[pred], features=feat.unsqueeze(1))
    feats, preds = [], []
    for i in range(n_experts):
        ds = None if drop_scales is None else drop_scales[i]
        f, z = expert_forward(sd, i, image, bn_mode, ds)
        feats.append(f)
        preds.append(z)
    x = torch.stack(feats, 1)                                                        # :400 / :369
    ro = dm_router(sd, x)
    r = gate_scores(sd, ro)
    if is_train:
        gate = torch.softmax(1.0 * r, dim=-1)                                        # :406,495-496 beta=1
        return dict(logits=combine_soft(preds, gate), index=gate, preds=preds, features=x, router_out=ro, scores=r)
    index = torch.max(r, -1)[1]                                                      # :376-377
    return dict(logits=combine_hard(preds, index), index=index, preds=preds, features=x, router_out=ro, scores=r)


# ----------------------------------------------------------------------------------------------
# CTC loss restated from scratch (il_modules/base.py:131; call sites il_modules/mrn.py:251-252,345-346)
# ----------------------------------------------------------------------------------------------

NEG_INF = float("-inf")


def ctc_nll_and_grad(logits: torch.Tensor, targets: torch.Tensor, target_lengths: torch.Tensor,
                     want_grad: bool = True):
    """log_softmax over C, then the CTC alpha-beta recursion with blank = 0 over all T frames
    (input_lengths = T for every sample, il_modules/mrn.py:250), `zero_infinity=True`.

    logits [B,T,C]; targets [B,S] (padded with 1); target_lengths [B].
    Returns nll [B] (0 where infeasible) and, if want_grad, dnll/dlogits [B,T,C] of the *per-sample* nll
    (softmax - occupancy; zero rows for infeasible samples)."""
    B, T, C = logits.shape
    dt = torch.float64
    lp = torch.log_softmax(logits.to(dt), dim=-1)
    nll = torch.zeros(B, dtype=dt)
    grad = torch.zeros(B, T, C, dtype=dt) if want_grad else None
    for b in range(B):
        L = int(target_lengths[b])
        S = 2 * L + 1
        ext = torch.zeros(S, dtype=torch.long)
        ext[1::2] = targets[b, :L].long()
        lpe = lp[b][:, ext]                                      # [T,S]
        can_skip = torch.zeros(S, dtype=torch.bool)
        if S > 2:
            can_skip[2:] = (ext[2:] != 0) & (ext[2:] != ext[:-2])
        alpha = torch.full((T, S), NEG_INF, dtype=dt)
        alpha[0, 0] = lpe[0, 0]
        if S > 1:
            alpha[0, 1] = lpe[0, 1]
        for t in range(1, T):
            a = alpha[t - 1]
            a1 = torch.cat([torch.full((1,), NEG_INF, dtype=dt), a])[:S]
            a2 = torch.cat([torch.full((2,), NEG_INF, dtype=dt), a])[:S]
            a2 = torch.where(can_skip, a2, torch.full_like(a2, NEG_INF))
            alpha[t] = torch.logsumexp(torch.stack([a, a1, a2]), 0) + lpe[t]
        tail = alpha[T - 1, S - 1:S] if S == 1 else alpha[T - 1, S - 2:S]
        ll = torch.logsumexp(tail, 0)
        if not torch.isfinite(ll):                               # zero_infinity
            continue
        nll[b] = -ll
        if not want_grad:
            continue
        beta = torch.full((T, S), NEG_INF, dtype=dt)
        beta[T - 1, S - 1] = lpe[T - 1, S - 1]
        if S > 1:
            beta[T - 1, S - 2] = lpe[T - 1, S - 2]
        for t in range(T - 2, -1, -1):
            bt = beta[t + 1]
            b1 = torch.cat([bt, torch.full((1,), NEG_INF, dtype=dt)])[1:S + 1]
            b2 = torch.cat([bt, torch.full((2,), NEG_INF, dtype=dt)])[2:S + 2]
            skip_from = torch.zeros(S, dtype=torch.bool)
            if S > 2:
                skip_from[:-2] = can_skip[2:]
            b2 = torch.where(skip_from, b2, torch.full_like(b2, NEG_INF))
            beta[t] = torch.logsumexp(torch.stack([bt, b1, b2]), 0) + lpe[t]
        occ = torch.exp(alpha + beta - lpe - ll)                 # [T,S] posterior state occupancy
        g = torch.exp(lp[b])                                     # softmax
        g.index_add_(1, ext, -occ)
        grad[b] = g
    return nll, grad


def ctc_loss_mean(logits, targets, target_lengths):
    """torch.nn.CTCLoss(reduction='mean', zero_infinity=True): mean_b(nll_b / max(len_b, 1))
    (SURVEY.md Appendix A.5)."""
    nll, _ = ctc_nll_and_grad(logits, targets, target_lengths, want_grad=False)
    return (nll / target_lengths.clamp(min=1).to(nll.dtype)).mean()


def ctc_brute_force(logits: torch.Tensor, target: Sequence[int]) -> float:
    """-log sum over all alignments (tiny T, C only) -- independent check of ctc_nll_and_grad."""
    import itertools
    T, C = logits.shape
    p = torch.softmax(logits.double(), -1)
    tot = 0.0
    for path in itertools.product(range(C), repeat=T):
        col, prev = [], None
        for s in path:
            if s != prev and s != 0:
                col.append(s)
            prev = s
        if col == list(target):
            pr = 1.0
            for t, s in enumerate(path):
                pr *= float(p[t, s])
            tot += pr
    return -math.log(tot) if tot > 0 else float("inf")


# ----------------------------------------------------------------------------------------------
# Stage-1 (router-training) loss and gradients  (il_modules/mrn.py:298-371)
# ----------------------------------------------------------------------------------------------

ROUTER_KEYS = ("route.weight", "route.bias", "channel_route.weight", "channel_route.bias",
               "dm_router.0.norm.weight", "dm_router.0.norm.bias",
               "dm_router.0.proj_1.weight", "dm_router.0.proj_1.bias",
               "dm_router.0.spatial_gating.norm.weight", "dm_router.0.spatial_gating.norm.bias",
               "dm_router.0.spatial_gating.proj.weight", "dm_router.0.spatial_gating.proj.bias",
               "dm_router.0.channel_gating.norm.weight", "dm_router.0.channel_gating.norm.bias",
               "dm_router.0.channel_gating.proj.weight", "dm_router.0.channel_gating.proj.bias",
               "dm_router.0.proj_2.weight", "dm_router.0.proj_2.bias",
               "dm_router.0.proj_3.weight", "dm_router.0.proj_3.bias")   # nn.Module.parameters() order, model.py:437-452


def stage1_loss_from_features(sd, features, preds, targets, target_lengths, domain, pi: float = 15.0):
    """Router-training objective given frozen expert outputs (il_modules/mrn.py:338-360):
    loss = pi * CTC(log_softmax(sum_i g_i pad_i)) + CrossEntropy(g, domain)  -- CE applied to the already
    softmaxed gate (reference quirk 2).  Differentiable w.r.t. the router parameters in `sd`."""
    ro = dm_router(sd, features)
    r = gate_scores(sd, ro)
    gate = torch.softmax(r, dim=-1)
    logits = combine_soft(preds, gate)
    lp = logits.log_softmax(2).permute(1, 0, 2)
    B, T = logits.shape[0], logits.shape[1]
    loss_clf = F.ctc_loss(lp, targets, torch.full((B,), T, dtype=torch.int32), target_lengths.to(torch.int32),
                          blank=0, reduction="mean", zero_infinity=True)
    taski_loss = F.cross_entropy(gate, domain)
    return pi * loss_clf + taski_loss, loss_clf, taski_loss, gate, logits, ro, r


def stage1_router_grads(sd, features, preds, targets, target_lengths, domain, pi: float = 15.0, dtype=torch.float64):
    """Autograd over the restatement: gradients of the stage-1 loss w.r.t. every router parameter."""
    sd2 = {k: v.detach().to(dtype) if v.is_floating_point() else v for k, v in sd.items()
           if k in ROUTER_KEYS}
    for k in ROUTER_KEYS:
        sd2[k].requires_grad_(True)
    out = stage1_loss_from_features(sd2, features.to(dtype), [p.to(dtype) for p in preds],
                                    targets, target_lengths, domain, pi)
    loss = out[0]
    grads = torch.autograd.grad(loss, [sd2[k] for k in ROUTER_KEYS])
    return dict(loss=loss.detach(), loss_clf=out[1].detach(), taski_loss=out[2].detach(), gate=out[3].detach(),
                logits=out[4].detach(), router_out=out[5].detach(), scores=out[6].detach(),
                grads={k: g for k, g in zip(ROUTER_KEYS, grads)})


def gate_grad_shortcut(logits, preds, gate, targets, target_lengths, pi: float = 15.0):
    """SURVEY.md Appendix A.5 identity used by the fused kernel: with G = d(pi*L_ctc)/dlogits,
    dL/dgate[b,i] = sum_{t,c} G[b,t,c] * pad_i[b,t,c].  Returns dL/dgate [B,I] (CTC part only)."""
    nll, g = ctc_nll_and_grad(logits, targets, target_lengths)
    B = logits.shape[0]
    scale = pi / (B * target_lengths.clamp(min=1).to(g.dtype))
    G = g * scale.view(-1, 1, 1)
    C = logits.shape[-1]
    return torch.stack([(G * pad_ones(p.to(g.dtype), C)).sum(dim=(1, 2)) for p in preds], 1)


def clip_and_adam(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: dict, lr: float,
                  max_norm: float = 5.0, betas=(0.9, 0.999), eps: float = 1e-8):
    """torch.nn.utils.clip_grad_norm_(.., 5) followed by torch.optim.Adam.step (il_modules/mrn.py:364-367),
    written out.  state: dict(step=int, m={k:..}, v={k:..}); updated in place.  Returns the pre-clip norm."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    coef = min(1.0, max_norm / (total + 1e-6))
    state["step"] += 1
    t = state["step"]
    for k, p in params.items():
        g = grads[k].to(p.dtype) * coef
        m = state["m"].setdefault(k, torch.zeros_like(p))
        v = state["v"].setdefault(k, torch.zeros_like(p))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1 = 1 - betas[0] ** t
        bc2 = 1 - betas[1] ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)
    return total


def one_cycle_lr(step: int, total_steps: int, max_lr: float, div_factor=20.0, final_div_factor=1000.0, pct_start=0.3):
    """torch.optim.lr_scheduler.OneCycleLR (cos anneal, two phases) as configured at il_modules/mrn.py:77-84;
    `step` = number of scheduler.step() calls made so far (lr used by optimizer step number step+1)."""
    initial = max_lr / div_factor
    min_lr = initial / final_div_factor
    end1 = float(pct_start * total_steps) - 1
    end2 = total_steps - 1

    def cos(a, b, pct):
        return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1)
    if step <= end1:
        return cos(initial, max_lr, step / end1)
    return cos(max_lr, min_lr, (step - end1) / (end2 - end1))


# ----------------------------------------------------------------------------------------------
# Greedy decode + confidence  (test.py:211-221,257; tools/utils.py:62-76)
# ----------------------------------------------------------------------------------------------

def greedy_decode(logits: torch.Tensor):
    """k[b,t] = argmax_c (first maximal index wins, as torch.max); collapse repeats, drop blank 0;
    confidence = prod_t max_c softmax(logits[b,t]) over ALL frames (cumprod(...)[-1]).

    Returns (raw ids [B,T] int64, list of compact id lists, confidence [B] in logits.dtype)."""
    _, k = logits.max(2)
    conf = torch.softmax(logits, dim=2).max(dim=2)[0].cumprod(dim=1)[:, -1]
    out = []
    for b in range(k.shape[0]):
        seq, prev = [], -1
        for t in range(k.shape[1]):
            v = int(k[b, t])
            if v != 0 and v != prev:
                seq.append(v)
            prev = v
        out.append(seq)
    return k, out, conf


def ctc_encode(words: Sequence[str], char_to_idx: Dict[str, int], batch_max_length: int = 25):
    """CTCLabelConverter.encode (tools/utils.py:35-60): pad with [PAD]=1, unknown -> [UNK]=2."""
    idx = torch.full((len(words), batch_max_length), 1, dtype=torch.long)
    lens = torch.zeros(len(words), dtype=torch.int32)
    for i, w in enumerate(words):
        ids = [char_to_idx.get(ch, 2) for ch in w]
        idx[i, :len(ids)] = torch.tensor(ids, dtype=torch.long)
        lens[i] = len(w)
    return idx, lens


# ----------------------------------------------------------------------------------------------
# One full router-training iteration on the CPU (bench.py cpu_baseline / --impl reference)
# ----------------------------------------------------------------------------------------------

def stage1_step_cpu(sd, n_experts, state, image, targets, target_lengths, domain, lr=5e-4, pi=15.0, bn_mode="batch",
                    drop_scales=None):
    """il_modules/mrn.py:329-371 end to end in fp32 on the host: frozen experts forward (no grad), router forward,
    loss = pi*CTC + CE, autograd backward through the router restatement, clip_grad_norm_(5), Adam.  Updates `sd` in
    place.  Returns (loss_clf, taski_loss)."""
    with torch.no_grad():
        feats, preds = [], []
        for i in range(n_experts):
            ds = None if drop_scales is None else drop_scales[i]
            f, z = expert_forward(sd, i, image, bn_mode, ds)
            feats.append(f)
            preds.append(z)
        x = torch.stack(feats, 1)
    r = stage1_router_grads(sd, x, preds, targets, target_lengths, domain, pi, dtype=torch.float32)
    params = {k: sd[k] for k in ROUTER_KEYS}
    clip_and_adam(params, r["grads"], state, lr)
    return float(r["loss_clf"]), float(r["taski_loss"])


# ----------------------------------------------------------------------------------------------
# Stage 0: training the newest expert end to end (il_modules/mrn.py:225-279)
# ----------------------------------------------------------------------------------------------

def expert_param_keys(sd, i: int) -> List[str]:
    """Trainable tensors of expert i that receive a gradient in stage 0: everything under model.{i}. except the
    BatchNorm buffers, the aliased Prediction.* (same object as fc, modules/model.py:181) and the three modules the SVTR
    forward never touches (modules/svtr.py:464-479: linear, last_conv, norm)."""
    p = f"model.{i}."
    dead = (p + "model.FeatureExtraction.ConvNet.linear.", p + "model.FeatureExtraction.ConvNet.last_conv.",
            p + "model.FeatureExtraction.ConvNet.norm.")
    return [k for k in sd if k.startswith(p) and not k.startswith(p + "Prediction.") and not k.startswith(dead)
            and "running_" not in k and "num_batches_tracked" not in k]


def stage0_loss_and_grads(sd, i: int, image, targets, target_lengths, bn_mode="batch", drop_scales=None,
                          dtype=torch.float64):
    """One stage-0 objective evaluation (il_modules/mrn.py:246-260): preds = model(image, cross=False)['logits'] is the
    LAST expert's "predict" (modules/model.py:351-353); loss = CTCLoss(mean, zero_infinity)(log_softmax(preds));
    gradients by autograd over this restatement.  drop_scales: None or [12,2,B] for this expert."""
    keys = expert_param_keys(sd, i)
    sd2 = {k: (v.detach().to(dtype) if v.is_floating_point() else v) for k, v in sd.items() if k.startswith(f"model.{i}.")}
    for k in keys:
        sd2[k].requires_grad_(True)
    _, logits = expert_forward(sd2, i, image.to(dtype), bn_mode, drop_scales)
    B, T = logits.shape[0], logits.shape[1]
    lp = logits.log_softmax(2).permute(1, 0, 2)
    loss = F.ctc_loss(lp, targets, torch.full((B,), T, dtype=torch.int32), target_lengths.to(torch.int32), blank=0,
                      reduction="mean", zero_infinity=True)
    grads = torch.autograd.grad(loss, [sd2[k] for k in keys], allow_unused=True)
    return dict(loss=loss.detach(), logits=logits.detach(),
                grads={k: (g if g is not None else torch.zeros_like(sd2[k])) for k, g in zip(keys, grads)})


def stage0_step_cpu(sd, i: int, state, image, targets, target_lengths, lr=5e-4, bn_mode="batch", drop_scales=None):
    """il_modules/mrn.py:236-267 on the host in fp32: forward, CTC, backward, clip_grad_norm_(5), Adam.  Updates the
    expert's entries of `sd` in place (BatchNorm running statistics are not tracked here).  Returns the loss."""
    r = stage0_loss_and_grads(sd, i, image, targets, target_lengths, bn_mode, drop_scales, dtype=torch.float32)
    params = {k: sd[k] for k in r["grads"]}
    clip_and_adam(params, r["grads"], state, lr)
    return float(r["loss"])
