"""Conv root representation for pixel observations (BASELINE config 5: 84x84x4 uint8 frames -> embedding 256).

The reference's `ResNetRepresentation` (muax/nn.py:313-331: x / 255 -> conv3x3/2 + relu -> 2 residual blocks ->
conv3x3/2 + relu -> 3 blocks -> avg-pool/2 -> 3 blocks -> avg-pool/2 -> `min_max_normalize2d`, residual block =
muax/nn.py:118-149 with LayerNorm over (H, W, C)) runs ONCE per act, at the root; everything inside the search loop
is the flat-embedding Dynamic / Prediction path of the CUDA engines.  So the torso is a plain torch / cuDNN module
(library code: no kernel claim) whose output — flattened, projected to `embedding_dim`, min-max normalised like
`Representation` (muax/nn.py:59-70) — is handed to the search as `root = (None, None, embedding)`.

`ResNetTorso` / `ResNetPrediction` / `ResNetDynamic` are the reference's all-conv network (muax/nn.py:313-395) as torch
modules: with them Prediction and Dynamic stay torch callables inside the simulation loop and the search runs through
the library's callback mode (tree kernels native, nets in torch): `create_resnet_muzero_network`.
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


def _min_max_normalize2d(x):
    """muax/nn.py:48-56 on NCHW tensors: per sample and channel over the spatial positions."""
    lo, hi = x.amin(dim=(2, 3), keepdim=True), x.amax(dim=(2, 3), keepdim=True)
    scale = hi - lo
    return (x - lo) / torch.where(scale < 1e-5, scale + 1e-5, scale)


class ResNetTorso(nn.Module):
    """The reference's `ResNetRepresentation` as is (muax/nn.py:313-331): frames [B, H, W, C] -> the spatial embedding,
    returned FLAT in NHWC order ([B, h * w * c], what `hk.Flatten` would give) so that it can live in the tree."""

    def __init__(self, input_channels=32, frame_channels=4, height=84, width=84):
        super().__init__()
        self.torso = ResNetRepresentation(8, input_channels, frame_channels, height, width).torso
        with torch.no_grad():
            c, h, w = self.torso(torch.zeros(1, frame_channels, height, width)).shape[1:]
        self.shape = (int(h), int(w), int(c))  # NHWC shape of one embedding

    @torch.no_grad()
    def forward(self, obs):
        x = obs.permute(0, 3, 1, 2).to(torch.float32, memory_format=torch.channels_last) / 255.0
        x = _min_max_normalize2d(self.torso(x))
        return x.permute(0, 2, 3, 1).reshape(x.shape[0], -1).contiguous()


def _head(cin, channels, n_convs, spatial, out_dim):
    layers, c = [], cin
    for _ in range(n_convs):
        layers += [nn.Conv2d(c, channels, 1, bias=False), nn.ReLU()]
        c = channels
    return nn.Sequential(*layers, nn.Flatten(), nn.Linear(channels * spatial, channels), nn.ReLU(),
                         nn.Linear(channels, out_dim))


class ResNetPrediction(nn.Module):
    """muax/nn.py:334-361: 1x1-conv value head (two convs) and policy head (one conv) over the spatial embedding."""

    def __init__(self, shape, num_actions, full_support_size, output_channels=16):
        super().__init__()
        self.shape = tuple(shape)
        h, w, c = self.shape
        self.v_func = _head(c, output_channels, 2, h * w, full_support_size)
        self.pi_func = _head(c, output_channels, 1, h * w, num_actions)

    @torch.no_grad()
    def forward(self, s):
        h, w, c = self.shape
        x = s.reshape(-1, h, w, c).permute(0, 3, 1, 2)
        return self.v_func(x), self.pi_func(x)


class ResNetDynamic(nn.Module):
    """muax/nn.py:364-395: the action as a constant plane a / num_actions appended to the embedding; reward head of
    two 1x1 convs + two linears; next state = 1x1 conv + 8 residual blocks, min-max normalised per channel."""

    def __init__(self, shape, num_actions, full_support_size, output_channels=64):
        super().__init__()
        self.shape, self.num_actions = tuple(shape), num_actions
        h, w, c = self.shape
        self.r_func = _head(c + 1, output_channels, 2, h * w, full_support_size)
        self.ns_func = nn.Sequential(nn.Conv2d(c + 1, output_channels, 1, bias=False), nn.ReLU(),
                                     *[_ResBlockV1(output_channels, output_channels) for _ in range(8)])
        if output_channels != c:
            raise ValueError("the next state must have the embedding's channel count (reference: 64)")

    @torch.no_grad()
    def forward(self, s, a):
        h, w, c = self.shape
        x = s.reshape(-1, h, w, c).permute(0, 3, 1, 2)
        plane = (a.to(torch.float32) / self.num_actions).view(-1, 1, 1, 1).expand(-1, 1, h, w)
        sa = torch.cat([x, plane], dim=1)
        ns = _min_max_normalize2d(self.ns_func(sa))
        return self.r_func(sa), ns.permute(0, 2, 3, 1).reshape(ns.shape[0], -1).contiguous()


def create_resnet_muzero_network(num_actions, full_support_size, input_channels=32, frame_channels=4, height=84,
                                 width=84, device="cuda"):
    """The reference's all-conv MuZero (`ResNetRepresentation` / `ResNetPrediction` / `ResNetDynamic`) as an
    `MZNetwork` of torch callables: `muax_b200.MuZero(network, ...)` then searches through the callback mode."""
    from .nn import MZNetwork
    torso = ResNetTorso(input_channels, frame_channels, height, width).to(device).eval()
    pred = ResNetPrediction(torso.shape, num_actions, full_support_size).to(device).eval()
    dyn = ResNetDynamic(torso.shape, num_actions, full_support_size, output_channels=torso.shape[2]).to(device).eval()
    return MZNetwork(torso, pred, dyn)
