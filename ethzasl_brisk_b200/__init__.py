"""B200-native BRISK hot path (pyramid -> AGAST/Harris scale-space detection ->
BRISK descriptors -> Hamming kNN) behind the reference's class surface.

Importing the package does not load CUDA; constructing a Context does and
fails loudly when libbrisk_b200.so or a GPU is missing (no CPU fallback).
"""
from . import build
from .api import (KP_DTYPE, STAGES, BriskDescriptorExtractor, BriskError, BriskFeature, BriskFeatureDetector, BruteForceMatcher,
                  Context, Hamming, HarrisFeatureDetector, HarrisScaleSpaceFeatureDetector, HarrisScoreCalculator, ScaleSpaceFeatureDetector, default_context,
                  detect_and_compute_batch, lib_path, load_library)
from .setio import read_pgm, read_set, write_pgm, write_set
from .synthetic import random_descriptors, synthetic_batch, synthetic_frame

__all__ = ["KP_DTYPE", "STAGES", "BriskDescriptorExtractor", "BriskError", "BriskFeature", "BriskFeatureDetector", "BruteForceMatcher",
           "HarrisFeatureDetector", "HarrisScaleSpaceFeatureDetector", "HarrisScoreCalculator", "ScaleSpaceFeatureDetector",
           "Context", "Hamming", "default_context", "detect_and_compute_batch", "lib_path", "load_library",
           "random_descriptors", "synthetic_batch", "synthetic_frame", "read_pgm", "read_set", "write_pgm", "write_set"]
