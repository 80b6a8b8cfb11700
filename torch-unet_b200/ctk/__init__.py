"""ctk -- B200-native hot path of CrosstalkPy behind the reference's own Python surface.

    from ctk import AdvancedRegressionModel, SimplifiedTwoBranchRegressionModel, accelerate
    from ctk import pearson_per_image, mse_loss, Adam

Everything here drives hand-written sm_100a kernels in libctk.so through the C ABI of include/ctk.h.
"""
from ._lib import CtkError, EXPORTED_SYMBOLS, LIB_PATH, load
from .engine import InferenceEngine
from .metrics import nmi_per_image, pearson_per_image, ssim_per_image, tile_metrics
from .models import (AdvancedRegressionModel, SimplifiedFeatureExtractionBranch, SimplifiedRegressionHead,
                     SimplifiedTwoBranchRegressionModel, accelerate, set_precision)
from .optim import Adam, MSELoss, mse_loss
from .schedulers import CosineWarmupLR
from .pipeline import DevicePrefetcher, HostScorer, prefetch_to_device
from . import io, parallel, synthetic
from .io import prepare_tiles

__all__ = ["CtkError", "EXPORTED_SYMBOLS", "LIB_PATH", "load", "InferenceEngine", "pearson_per_image", "tile_metrics", "nmi_per_image", "ssim_per_image",
           "AdvancedRegressionModel", "SimplifiedFeatureExtractionBranch", "SimplifiedRegressionHead",
           "SimplifiedTwoBranchRegressionModel", "accelerate", "set_precision", "Adam", "MSELoss", "mse_loss", "CosineWarmupLR", "HostScorer", "DevicePrefetcher", "prefetch_to_device", "parallel", "io", "prepare_tiles"]
