"""irp_b200 — B200-native degradation analysis + preprocess for image-restoration-platform.

Only the hot path lives here: the CUDA kernels + C ABI (csrc/, libirp_b200.so) and
the host-side mirrors of the two reference interfaces that sit on it:
  classifier.ClassifierService   <- server-node/src/services/classifier.js
  preprocess.preprocess_image    <- server-node/src/middleware/imagePreprocess.js
"""
from .engine import DeviceImage, Engine, IrpError, SCORE_KEYS  # noqa: F401
from .classifier import ClassifierService, createClassifierService, create_classifier_service, DEGRADATION_TYPES  # noqa: F401
from .preprocess import Problem, preprocess_image, preprocessImage  # noqa: F401

__all__ = ["Engine", "DeviceImage", "IrpError", "SCORE_KEYS", "ClassifierService", "createClassifierService",
           "create_classifier_service", "DEGRADATION_TYPES", "Problem", "preprocess_image", "preprocessImage"]
