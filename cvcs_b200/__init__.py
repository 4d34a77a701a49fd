"""cvcs_b200 — B200-native (sm_100a) per-pixel hot path of the theElandor/CVCS segmentation
pipeline: tiling/normalisation, fused softmax-CE fwd+bwd + argmax + confusion matrix, behind the
reference's own Python call signatures and a plain C-ABI (include/cvcs_b200.h).

    from cvcs_b200.loss import load_loss, FusedCrossEntropyLoss      # utils.load_loss
    from cvcs_b200.metrics import eval_model, print_metrics          # utils.eval_model / print_metrics
    from cvcs_b200.dataset import Loader                             # dataset.Loader
    from cvcs_b200.inference import GID15, inference_scene           # dataset.GID15, utils.inference + re-assembly
    from cvcs_b200 import shard                                      # tile sharding + the three tiny collectives

Importing the package does not load the CUDA library; the first use of cvcs_b200._lib does, and it
raises if libcvcs_b200.so has not been built (`python -m cvcs_b200.build`).  No CPU fallback.
"""
__version__ = "0.1.0"
