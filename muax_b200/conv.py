"""Conv root representation for pixel observations (BASELINE config 5: 84x84x4 uint8 frames -> embedding 256).

The reference's `ResNetRepresentation` (muax/nn.py:313-331: x / 255 -> conv3x3/2 + relu -> 2 residual blocks ->
conv3x3/2 + relu -> 3 blocks -> avg-pool/2 -> 3 blocks -> avg-pool/2 -> `min_max_normalize2d`, residual block =
muax/nn.py:118-149 with LayerNorm over (H, W, C)) runs ONCE per act, at the root; everything inside the search loop
is the flat-embedding Dynamic / Prediction path of the CUDA engines.  So the torso is a plain torch / cuDNN module
(library code: no kernel claim) whose output — flattened, projected to `embedding_dim`, min-max normalised like
`Representation` (muax/nn.py:59-70) — is handed to the search as `root = (None, None, embedding)`.
"""
import torch
from torch import nn


class _ResBlockV1(nn.Module):  # muax/nn.py:118-149 (use_projection=True everywhere in ResNetRepresentation)
    def __init__(self, cin, channels):
        super().__init__()
        self.proj_conv = nn.Conv2d(cin, channels, 3, 1, 1, bias=False)
        self.proj_ln = nn.GroupNorm(1, channels)  # LayerNorm over (H, W, C) with per-channel scale / offset
        self.conv_0 = nn.Conv2d(cin, channels, 3, 1, 1, bias=False)
        self.ln_0 = nn.GroupNorm(1, channels)
        self.conv_1 = nn.Conv2d(channels, channels, 3, 1, 1, bias=False)
        self.ln_1 = nn.GroupNorm(1, channels)

    def forward(self, x):
        shortcut = self.proj_ln(self.proj_conv(x))
        out = self.ln_1(self.conv_1(torch.relu(self.ln_0(self.conv_0(x)))))
        return torch.relu(shortcut + out)


class ResNetRepresentation(nn.Module):
    """obs uint8 / float [B, H, W, C] (the reference's NHWC frames) -> embedding float32 [B, embedding_dim]."""

    def __init__(self, embedding_dim=256, input_channels=32, frame_channels=4, height=84, width=84):
        super().__init__()
        c = input_channels
        layers = [nn.Conv2d(frame_channels, c, 3, 2, 1, bias=False), nn.ReLU()]
        layers += [_ResBlockV1(c, c) for _ in range(2)]
        layers += [nn.Conv2d(c, 2 * c, 3, 2, 1, bias=False), nn.ReLU()]
        layers += [_ResBlockV1(2 * c, 2 * c) for _ in range(3)]
        layers += [nn.AvgPool2d(3, 2, 1)]
        layers += [_ResBlockV1(2 * c, 2 * c) for _ in range(3)]
        layers += [nn.AvgPool2d(3, 2, 1)]
        self.torso = nn.Sequential(*layers)
        with torch.no_grad():
            n = self.torso(torch.zeros(1, frame_channels, height, width)).numel()
        self.project = nn.Linear(n, embedding_dim)
        self.embedding_dim = embedding_dim
        self._torso16 = None  # bf16 copy of the torso, made on first use (throughput mode); stale after a weight update

    supports_bf16 = True  # MuZero._plan passes bf16=True in the throughput mode (precision="bf16")

    @torch.no_grad()
    def forward(self, obs, bf16=False):
        """uint8 frames stay uint8 until they are on the device (a quarter of the float32 H2D bytes); bf16=True runs
        the convolutions under autocast (tensor cores; the normalisations stay fp32)."""
        x = obs.permute(0, 3, 1, 2).to(torch.float32, memory_format=torch.channels_last) / 255.0
        if bf16 and x.is_cuda:
            # the torso is bandwidth-bound (24 convolutions + 24 normalisations over 42 x 42 x 32 activations): a bf16
            # copy of its weights and bf16 activations halve the bytes (statistics are still accumulated in fp32)
            if self._torso16 is None:
                import copy
                self._torso16 = copy.deepcopy(self.torso).to(torch.bfloat16)
            x = self._torso16(x.to(torch.bfloat16)).to(torch.float32)
        else:
            x = self.torso(x)
        # min_max_normalize2d (muax/nn.py:48-56): per sample and channel over the spatial positions
        lo, hi = x.amin(dim=(2, 3), keepdim=True), x.amax(dim=(2, 3), keepdim=True)
        scale = hi - lo
        x = (x - lo) / torch.where(scale < 1e-5, scale + 1e-5, scale)
        s = self.project(x.flatten(1))
        lo, hi = s.amin(dim=1, keepdim=True), s.amax(dim=1, keepdim=True)  # min_max_normalize (muax/nn.py:37-44)
        scale = hi - lo
        return ((s - lo) / torch.where(scale < 1e-5, scale + 1e-5, scale)).contiguous()
