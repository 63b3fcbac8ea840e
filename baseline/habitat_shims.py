"""HARNESS ONLY (tests + bench.py's `policy_forward` record) -- stand-ins for the third-party packages the reference's
policy imports but this image does not have: `gym`, `habitat` and `habitat_baselines` (habitat-lab v0.1.5, pinned in the
reference's SETUP.md).  They exist so that the UNMODIFIED reference files under baseline/_ref/ (MGMapNet,
mg_map_policy.py; BasePolicy, policy.py; the UNet / map / depth / instruction encoders) can be constructed with random
weights and stepped, which is what BASELINE.json configs[4] asks for: "full CMA policy forward ... with the new map
kernels swapped in, batch 64".

Nothing here is on the product path and nothing here restates the map update.  The stand-ins follow the published
interfaces of those packages (constructor arguments, attribute names, tensor shapes) -- enough for the reference code to
run -- and are written from that interface description, not from their sources:

  gym.Space / gym.spaces.Box / gym.spaces.Dict          shape containers
  habitat.Config                                        attribute-style nested config
  habitat_baselines.rl.ppo.policy.Net / CriticHead      abstract base / linear value head
  habitat_baselines.rl.models.rnn_state_encoder.RNNStateEncoder   masked single-step GRU/LSTM
  habitat_baselines.rl.ddppo.policy.resnet.resnet50 + resnet_policy.ResNetEncoder   GroupNorm ResNet-50, depth 256x256
        -> avg-pool 2 -> backbone (baseplanes 32) -> 3x3 compression conv: output_shape (128, 4, 4)
  habitat_baselines.common.utils.Flatten
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------ gym
class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = tuple(shape) if shape is not None else None
        self.dtype = dtype


class Box(Space):
    def __init__(self, low, high, shape, dtype="float32"):
        super().__init__(shape, dtype)
        self.low, self.high = low, high


class Dict(Space):
    def __init__(self, spaces):
        super().__init__(None, None)
        self.spaces = dict(spaces)


# ------------------------------------------------------------------ habitat.Config
class Config(dict):
    """Nested attribute-style config (what the reference reads: `cfg.A.b`, `'x' in cfg.A.list`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def of(d):
        c = Config()
        for k, v in d.items():
            c[k] = Config.of(v) if isinstance(v, dict) else v
        return c


# ------------------------------------------------------------------ habitat_baselines
class Net(nn.Module):
    """Abstract policy trunk: subclasses provide forward / output_size / num_recurrent_layers / is_blind."""


class CriticHead(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.fc = nn.Linear(input_size, 1)
        nn.init.orthogonal_(self.fc.weight)
        nn.init.constant_(self.fc.bias, 0)

    def forward(self, x):
        return self.fc(x)


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.size(0), -1)


class RNNStateEncoder(nn.Module):
    """One recurrent step per call when the input has one row per env: hidden state zeroed where masks == 0,
    hidden layout [num_recurrent_layers, bs, hidden] (LSTM: h and c stacked along dim 0)."""

    def __init__(self, input_size, hidden_size, num_layers=1, rnn_type="GRU"):
        super().__init__()
        self._num_recurrent_layers = num_layers
        self._rnn_type = rnn_type
        self.rnn = getattr(nn, rnn_type)(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers)
        for name, p in self.rnn.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(p)
            elif "bias" in name:
                nn.init.constant_(p, 0)

    @property
    def num_recurrent_layers(self):
        return self._num_recurrent_layers * (2 if "LSTM" in self._rnn_type else 1)

    def forward(self, x, hidden_states, masks):
        if x.size(0) != hidden_states.size(1):
            raise NotImplementedError("stand-in supports one step per env per call (what the policy forward uses)")
        h = hidden_states * masks.unsqueeze(0)
        if "LSTM" in self._rnn_type:
            h = tuple(t.contiguous() for t in torch.chunk(h, 2, 0))
        else:
            h = h.contiguous()
        y, h = self.rnn(x.unsqueeze(0), h)
        if "LSTM" in self._rnn_type:
            h = torch.cat(h, 0)
        return y.squeeze(0), h


def _gn(ngroups, ch):
    return nn.GroupNorm(ngroups, ch)


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, ngroups, stride):
        super().__init__()
        out = planes * self.expansion
        self.convs = nn.Sequential(
            nn.Conv2d(inplanes, planes, 1, bias=False), _gn(ngroups, planes), nn.ReLU(True),
            nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False), _gn(ngroups, planes), nn.ReLU(True),
            nn.Conv2d(planes, out, 1, bias=False), _gn(ngroups, out))
        self.down = None
        if stride != 1 or inplanes != out:
            self.down = nn.Sequential(nn.Conv2d(inplanes, out, 1, stride=stride, bias=False), _gn(ngroups, out))

    def forward(self, x):
        r = x if self.down is None else self.down(x)
        return F.relu(self.convs(x) + r, True)


class _ResNet(nn.Module):
    def __init__(self, in_channels, base_planes, ngroups, layers):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_channels, base_planes, 7, stride=2, padding=3, bias=False),
                                  _gn(ngroups, base_planes), nn.ReLU(True), nn.MaxPool2d(3, 2, 1))
        blocks, inplanes = [], base_planes
        for i, n in enumerate(layers):
            planes = base_planes * (2 ** i)
            for k in range(n):
                blocks.append(_Bottleneck(inplanes, planes, ngroups, stride=(1 if i == 0 else 2) if k == 0 else 1))
                inplanes = planes * _Bottleneck.expansion
        self.blocks = nn.Sequential(*blocks)
        self.final_channels = inplanes
        self.final_spatial_compress = 1.0 / 32

    def forward(self, x):
        return self.blocks(self.stem(x))


def resnet50(in_channels, base_planes, ngroups):
    return _ResNet(in_channels, base_planes, ngroups, [3, 4, 6, 3])


class ResNetEncoder(nn.Module):
    def __init__(self, observation_space, baseplanes=32, ngroups=32, spatial_size=128, make_backbone=None,
                 normalize_visual_inputs=False, obs_transform=None):
        super().__init__()
        depth = observation_space.spaces["depth"]
        spatial = depth.shape[0] // 2
        self.backbone = make_backbone(depth.shape[2], baseplanes, ngroups)
        final_spatial = int(-(-spatial * self.backbone.final_spatial_compress // 1))
        n_comp = int(round(2048 / (final_spatial ** 2)))
        self.compression = nn.Sequential(nn.Conv2d(self.backbone.final_channels, n_comp, 3, padding=1, bias=False),
                                         nn.GroupNorm(1, n_comp), nn.ReLU(True))
        self.output_shape = (n_comp, final_spatial, final_spatial)

    def forward(self, observations):
        x = observations["depth"].permute(0, 3, 1, 2)
        x = F.avg_pool2d(x, 2)
        return self.compression(self.backbone(x))


# ------------------------------------------------------------------ installation
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-ins under the third-party module names (idempotent; never shadows a real install)."""
    for probe in ("gym", "habitat", "habitat_baselines"):
        if probe in sys.modules and not getattr(sys.modules[probe], "_wsmg_standin", False):
            continue
        if probe not in sys.modules and importlib.util.find_spec(probe) is not None:
            continue
        if probe == "gym":
            sp = _module("gym.spaces", Space=Space, Box=Box, Dict=Dict)
            _module("gym", Space=Space, spaces=sp, _wsmg_standin=True)
        elif probe == "habitat":
            _module("habitat", Config=Config, _wsmg_standin=True)
        else:
            resnet = _module("habitat_baselines.rl.ddppo.policy.resnet", resnet50=resnet50)
            rp = _module("habitat_baselines.rl.ddppo.policy.resnet_policy", ResNetEncoder=ResNetEncoder)
            pol = _module("habitat_baselines.rl.ddppo.policy", resnet=resnet, resnet_policy=rp)
            ddppo = _module("habitat_baselines.rl.ddppo", policy=pol)
            ppo_pol = _module("habitat_baselines.rl.ppo.policy", Net=Net, CriticHead=CriticHead)
            ppo = _module("habitat_baselines.rl.ppo", policy=ppo_pol)
            rse = _module("habitat_baselines.rl.models.rnn_state_encoder", RNNStateEncoder=RNNStateEncoder)
            models = _module("habitat_baselines.rl.models", rnn_state_encoder=rse)
            rl = _module("habitat_baselines.rl", ddppo=ddppo, ppo=ppo, models=models)
            utils = _module("habitat_baselines.common.utils", Flatten=Flatten)
            common = _module("habitat_baselines.common", utils=utils)
            _module("habitat_baselines", rl=rl, common=common, _wsmg_standin=True)
