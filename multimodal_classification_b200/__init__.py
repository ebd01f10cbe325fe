"""B200-native (sm_100a) ViLBERT hot path: two-stream encoder fwd/bwd + ResNet-152 RoI feature stage.

Drop-in for the reference's ``models/vilbert_facebook_arch.py`` / ``feature_extractors/resnet152_roi.py``.
Host code is Python; every device op is a hand-written CUDA kernel reached through the C ABI in
``include/vilbert_b200.h`` (``libvilbert_b200.so``).  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
