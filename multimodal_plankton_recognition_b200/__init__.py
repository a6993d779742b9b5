"""B200-native (sm_100a) cross-modal similarity hot path of
imveikka/multimodal_plankton_recognition: the fused similarity + InfoNCE coordination loss
(``CLIPLoss``, drop-in for reference src/coordination.py:17-47) and euclidean/cosine top-k retrieval
with the inverse-distance k-NN vote (``ANNClassifier``, drop-in for reference src/ann.py:6-34).

Also on the same kernels: ``CLIPPlus``, ``SigLIPLoss``, ``SigLIPPlus`` (reference src/coordination.py:50-112).
Submodules: ``dist`` (row-sharded loss, gallery-sharded retrieval), ``prefetch`` (host -> HBM staging under the
running step), ``harness`` (the reference drivers' few-shot benchmark), ``ops`` (tensor-level access to the C calls).

All arithmetic runs in ``libplk.so`` (hand-written CUDA, C ABI in ``include/plk.h``); there is no CPU path.
"""
from .coordination import CLIPLoss, CLIPPlus, SigLIPLoss, SigLIPPlus  # noqa: F401
from .ann import ANNClassifier  # noqa: F401

__all__ = ["CLIPLoss", "CLIPPlus", "SigLIPLoss", "SigLIPPlus", "ANNClassifier"]
