"""Host-side mirror of the reference's model surface.

Same class names, constructor signatures, sub-module names/indices and therefore the same ``state_dict``
keys, ``str(model)`` and ``.pth`` compatibility as /root/reference/regression_model.py and
two_branch_regression.py -- but ``forward`` runs the hand-written sm_100a kernels of libctk instead of
ATen/cuDNN.  ``accelerate(model)`` does the same to an instance built by the reference's own classes, which
is how train_model.py / test-cross-talk-model.py pick the path up without a text change.

There is no CPU fallback: calling these models with CPU tensors raises.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from . import _lib
from .engine import InferenceEngine


def _block(cin: int, cout: int):
    # Conv3x3(p=1) -> BatchNorm2d -> LeakyReLU(0.01) -> MaxPool2d(2,2): the unit both reference models repeat
    return [nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU(0.01),
            nn.MaxPool2d(kernel_size=2, stride=2)]


def _head(in_features: int, p_drop: float, sigmoid: bool) -> nn.Sequential:
    layers = [nn.Flatten(),
              nn.Linear(in_features, 512), nn.BatchNorm1d(512), nn.LeakyReLU(0.01), nn.Dropout(p_drop),
              nn.Linear(512, 128), nn.BatchNorm1d(128), nn.LeakyReLU(0.01), nn.Dropout(p_drop),
              nn.Linear(128, 1)]
    if sigmoid:
        layers.append(nn.Sigmoid())
    return nn.Sequential(*layers)


def get_engine(model) -> InferenceEngine:
    """The (lazily created) libctk inference engine bound to ``model``."""
    eng = model.__dict__.get("_ctk_engine")
    if eng is None:
        eng = InferenceEngine(model, conv_flags=model.__dict__.get("_ctk_conv_flags", 0),
                              precision=model.__dict__.get("_ctk_precision", "bf16"))
        model.__dict__["_ctk_engine"] = eng
    return eng


def set_precision(model, precision: str):
    """Select the arithmetic of both paths.  "bf16" (default): bf16 operands, fp32 accumulation on tcgen05 (north_star bound
    1e-3 on the score).  "fp32": eval -- every operand carried as a bf16 (hi, lo) pair, three MMAs per product, one
    accumulator per K chunk (bound 1e-5); train -- the reference's own float32 arithmetic on the CUDA cores
    (ctk.train_f32.TrainEngineF32), for gradient / loss-curve parity runs.  Call it before parallel.attach()."""
    if precision not in ("bf16", "fp32"):
        raise _lib.CtkError("precision must be 'bf16' or 'fp32'")
    model.__dict__["_ctk_precision"] = precision
    model.__dict__.pop("_ctk_engine", None)
    model.__dict__.pop("_ctk_train_engine", None)
    return model


def get_train_engine(model):
    """The (lazily created) libctk training engine bound to ``model``."""
    eng = model.__dict__.get("_ctk_train_engine")
    if eng is None:
        if model.__dict__.get("_ctk_precision", "bf16") == "fp32":
            from .train_f32 import TrainEngineF32
            eng = TrainEngineF32(model)
        else:
            from .train import TrainEngine
            eng = TrainEngine(model)
        model.__dict__["_ctk_train_engine"] = eng
    return eng


def _ctk_forward(self, x):
    """forward() shared by the mirrored classes and by accelerate()d reference instances.

    train(): batch-statistics BatchNorm, Dropout, and an autograd edge whose backward runs the ctk dgrad / wgrad /
    BN-backward kernels.  eval(): folded-BN inference path.
    """
    if self.training:
        from .train import train_forward
        return train_forward(get_train_engine(self), x)
    return get_engine(self).forward(x)


class AdvancedRegressionModel(nn.Module):
    """Single-branch crosstalk regressor -- mirrors regression_model.py:5-61."""

    def __init__(self, input_channels=2, initial_filters=64, num_conv_blocks=5):
        super().__init__()
        layers, cin, cout = [], input_channels, initial_filters
        for i in range(num_conv_blocks):
            layers += _block(cin, cout)
            cin, cout = cout, min(cout * 2, 512)                    # regression_model.py:22 (cap at 512)
        self.conv_layers = nn.Sequential(*layers)
        # regression_model.py:31,52-56: the reference sizes the head with a train-mode dummy pass on zeros, which
        # also advances every conv-stack BN once.  Replay it so fresh models start identically.
        with torch.no_grad():
            feat = self.conv_layers(torch.zeros(1, input_channels, 256, 256))
        self.fc_layers = _head(int(feat[0].numel()), 0.1, sigmoid=False)

    forward = _ctk_forward


class SimplifiedFeatureExtractionBranch(nn.Module):
    """Mirrors two_branch_regression.py:5-35."""

    def __init__(self, in_channels=1, initial_filters=64):
        super().__init__()
        f = initial_filters
        self.conv_blocks = nn.Sequential(*(_block(in_channels, f) + _block(f, 2 * f) + _block(2 * f, 4 * f) +
                                           _block(4 * f, 8 * f)))

    def forward(self, x):
        raise _lib.CtkError("branches are executed by the parent model's fused ctk forward")


class SimplifiedRegressionHead(nn.Module):
    """Mirrors two_branch_regression.py:37-57."""

    def __init__(self, input_feature_size):
        super().__init__()
        self.fc_layers = _head(input_feature_size, 0.5, sigmoid=True)

    def forward(self, x):
        raise _lib.CtkError("the head is executed by the parent model's fused ctk forward")


class SimplifiedTwoBranchRegressionModel(nn.Module):
    """Double-branch crosstalk regressor -- mirrors two_branch_regression.py:59-100."""

    def __init__(self, initial_filters_per_branch=16, input_image_size=(256, 256)):
        super().__init__()
        self.bleed_branch = SimplifiedFeatureExtractionBranch(1, initial_filters_per_branch)
        self.source_branch = SimplifiedFeatureExtractionBranch(1, initial_filters_per_branch)
        h, w = input_image_size
        feat = (initial_filters_per_branch * 8 * 2) * (h // 16) * (w // 16)   # two_branch_regression.py:77-80
        self.regression_head = SimplifiedRegressionHead(feat)

    forward = _ctk_forward


def accelerate(model: nn.Module, conv_flags: int = 0, precision: str = "bf16") -> nn.Module:
    """Route ``model(inputs)`` of a reference-built instance through libctk, in place.

    Parameters, buffers, ``state_dict()``, ``str(model)`` and optimizers attached to the parameters are
    untouched; only the instance's ``forward`` is rebound.
    """
    InferenceEngine(model)            # validates the architecture now, loudly
    model.__dict__["_ctk_conv_flags"] = conv_flags
    set_precision(model, precision)
    model.forward = types.MethodType(_ctk_forward, model)
    return model
