"""peanut_b200: B200-native (sm_100a) implementation of PEANUT's per-step perception hot path.

Public surface mirrors the reference's call sites:
  peanut_b200.prediction.PEANUT_Prediction_Model / run_inference / init_segmentor   (stage C)
  peanut_b200.mapping.Semantic_Mapping                                               (stage B)
  peanut_b200.segmentation.SemanticPredMaskRCNN                                      (stage A)
All compute runs in libpeanut_b200.so (hand-written CUDA); importing the package does not load it,
constructing any of the classes does and raises if it is missing.
"""
__version__ = "0.1.0"
