"""Parameter holders of the SVTR backbone with the reference's module tree and state_dict keys
(modules/svtr.py:315-479; modules/feature_extraction.py:724-732).  They carry weights only: the forward pass is
the grouped CUDA path (mrn_b200.ops.svtr_experts_forward), never torch ops."""
import numpy as np
import torch
import torch.nn as nn

EMBED_DIM = (64, 128, 256)
DEPTH = (3, 6, 3)
NUM_HEADS = (2, 4, 8)


class _Holder(nn.Module):
    def forward(self, *a, **k):       # pragma: no cover
        raise RuntimeError("mrn_b200 parameter holder: compute runs in the grouped CUDA path, not per module")


class Attention(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class Mlp(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, dim * 4)
        self.fc2 = nn.Linear(dim * 4, dim)


class Block(_Holder):
    def __init__(self, dim, drop_path):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.mixer = Attention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim)
        self.drop_path_rate = float(drop_path)


class PatchEmbed(_Holder):
    def __init__(self, in_channels, embed_dim):
        super().__init__()
        self.proj = nn.Sequential(
            nn.Conv2d(in_channels, embed_dim // 2, 3, 2, 1), nn.BatchNorm2d(embed_dim // 2), nn.GELU(),
            nn.Conv2d(embed_dim // 2, embed_dim, 3, 2, 1), nn.BatchNorm2d(embed_dim), nn.GELU())


class SubSample(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=3, stride=[2, 1], padding=1)
        self.norm = nn.LayerNorm(cout)


class SVTR(_Holder):
    """Same constructor defaults as modules/svtr.py:317-343 for img 32x256; only in/out channels vary."""

    def __init__(self, in_channels=4, out_channels=512, drop_path_rate=0.1):
        super().__init__()
        self.patch_embed = PatchEmbed(in_channels, EMBED_DIM[0])
        self.pos_embed = nn.Parameter(torch.zeros(1, 512, EMBED_DIM[0]))
        dpr = np.linspace(0, drop_path_rate, sum(DEPTH))
        k = 0
        for s in range(3):
            blocks = nn.ModuleList([Block(EMBED_DIM[s], dpr[k + j]) for j in range(DEPTH[s])])
            k += DEPTH[s]
            setattr(self, f"blocks{s + 1}", blocks)
            setattr(self, f"sub_sample{s + 1}", SubSample(EMBED_DIM[s], EMBED_DIM[s + 1] if s < 2 else out_channels))
        # constructed but never used by the reference forward (modules/svtr.py:464-479); kept for strict state_dict loads
        self.linear = nn.Linear(384, 512)
        self.last_conv = nn.Conv2d(EMBED_DIM[2], out_channels, 1, 1, 0, bias=False)
        self.norm = nn.LayerNorm(EMBED_DIM[-1], eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        # modules/svtr.py:488-498 (incl. the quirk: LayerNorm *bias* ends at 1.0, weight stays at its default 1.0)
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 1.0)
        elif isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_in")

    def drop_path_rates(self):
        return [b.drop_path_rate for s in range(3) for b in getattr(self, f"blocks{s + 1}")]


class SVTR_FeatureExtractor(_Holder):
    def __init__(self, input_channel, output_channel=512):
        super().__init__()
        self.ConvNet = SVTR(in_channels=input_channel, out_channels=output_channel)
