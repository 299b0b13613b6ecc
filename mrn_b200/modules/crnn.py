"""Parameter holders of the CRNN expert (VGG feature extractor + BidirectionalLSTM) with the reference's module
tree, initialisation and state_dict keys (modules/feature_extraction.py:8-47, modules/sequence_modeling.py:4-22).
They carry weights only: the forward pass is the grouped CUDA path (mrn_b200.ops.crnn_experts_forward)."""
import torch.nn as nn


class _Holder(nn.Module):
    def forward(self, *a, **k):       # pragma: no cover
        raise RuntimeError("mrn_b200 parameter holder: compute runs in the grouped CUDA path, not per module")


class VGG_FeatureExtractor(_Holder):
    """Same nn.Sequential indices as the reference (0,3,6,8,11,12,14,15,18 carry parameters)."""

    def __init__(self, input_channel, output_channel=512):
        super().__init__()
        oc = [output_channel // 8, output_channel // 4, output_channel // 2, output_channel]
        self.output_channel = oc
        self.ConvNet = nn.Sequential(
            nn.Conv2d(input_channel, oc[0], 3, 1, 1), nn.ReLU(True), nn.MaxPool2d(2, 2),
            nn.Conv2d(oc[0], oc[1], 3, 1, 1), nn.ReLU(True), nn.MaxPool2d(2, 2),
            nn.Conv2d(oc[1], oc[2], 3, 1, 1), nn.ReLU(True),
            nn.Conv2d(oc[2], oc[2], 3, 1, 1), nn.ReLU(True), nn.MaxPool2d((2, 1), (2, 1)),
            nn.Conv2d(oc[2], oc[3], 3, 1, 1, bias=False), nn.BatchNorm2d(oc[3]), nn.ReLU(True),
            nn.Conv2d(oc[3], oc[3], 3, 1, 1, bias=False), nn.BatchNorm2d(oc[3]), nn.ReLU(True),
            nn.MaxPool2d((2, 1), (2, 1)),
            nn.Conv2d(oc[3], oc[3], 2, 1, 0), nn.ReLU(True))


class BidirectionalLSTM(_Holder):
    def __init__(self, input_size, hidden_size, output_size):
        super().__init__()
        self.rnn = nn.LSTM(input_size, hidden_size, bidirectional=True, batch_first=True)
        self.linear = nn.Linear(hidden_size * 2, output_size)
