"""Host-side mirror of the reference interface (no GPU needed)."""
import numpy as np
import pytest
import torch

from oracle import ann as oann


def test_cliploss_surface_matches_reference():
    from multimodal_plankton_recognition_b200 import CLIPLoss
    mod = CLIPLoss()                      # reference src/coordination.py:21-23
    assert list(mod.state_dict().keys()) == ["logit_scale"]
    assert mod.logit_scale.shape == torch.Size([]) and float(mod.logit_scale) == 1.0
    assert mod.logit_scale.dtype == torch.float32 and mod.logit_scale.requires_grad
    CLIPLoss(bias=True)                   # ctor argument accepted and unused, as in the reference
    x = torch.randn(6, 4)
    with pytest.raises(AssertionError, match="Batch size must be divisible by number of buckets!"):
        mod(image_emb=x, profile_emb=x, buckets=4)
    with pytest.raises(ValueError):
        CLIPLoss(precision="fp8")
    assert CLIPLoss(precision="fp16").precision == "fp16"


def test_sigliploss_surface_matches_reference():
    from multimodal_plankton_recognition_b200 import SigLIPLoss, SigLIPPlus
    mod = SigLIPLoss()                    # reference src/coordination.py:70-73
    assert sorted(mod.state_dict().keys()) == ["bias", "logit_scale"]
    assert float(mod.logit_scale) == 1.0 and float(mod.bias) == -10.0
    assert mod.bias.shape == torch.Size([]) and mod.bias.dtype == torch.float32 and mod.bias.requires_grad
    x = torch.randn(6, 4)
    with pytest.raises(AssertionError, match="Batch size must be divisible by number of buckets!"):
        mod(image_emb=x, profile_emb=x, buckets=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(image_emb=x, profile_emb=x, buckets=1)
    plus = SigLIPPlus(beta=0.5)           # reference src/coordination.py:98-104
    assert sorted(plus.state_dict().keys()) == ["siglip.bias", "siglip.logit_scale"] and plus.beta == 0.5


def test_no_cpu_fallback():
    from multimodal_plankton_recognition_b200 import CLIPLoss
    x = torch.randn(8, 4, requires_grad=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CLIPLoss()(image_emb=x, profile_emb=x, buckets=1)
    if not torch.cuda.is_available():
        from multimodal_plankton_recognition_b200 import ANNClassifier
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ANNClassifier(np.eye(4, dtype=np.float32), np.arange(4))


def test_get_weights_matches_reference_rule():
    from multimodal_plankton_recognition_b200.ann import ANNClassifier
    d = np.array([[0.0, 0.5, 0.0], [0.25, 0.5, 1.0]], dtype=np.float32)
    w = ANNClassifier._get_weights(None, d.copy())
    np.testing.assert_array_equal(w, oann.inverse_distance_weights(d))
    assert w.dtype == np.float32


def test_product_never_imports_oracle():
    import os
    import re
    from conftest import ROOT
    pkg = os.path.join(ROOT, "multimodal_plankton_recognition_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_prefetcher_argument_validation():
    """prefetch.HostPairPrefetcher rejects an unusable configuration before touching CUDA."""
    from multimodal_plankton_recognition_b200.prefetch import HostPairPrefetcher
    with pytest.raises(ValueError, match="depth"):
        HostPairPrefetcher(iter(()), "cuda:0", depth=1)
    with pytest.raises(ValueError, match="CUDA"):
        HostPairPrefetcher(iter(()), "cpu", depth=2)


def test_sharded_module_keeps_nccl_path_without_cuda():
    """CLIPLoss._peer_scalars only turns the peer-memory exchange on for CUDA tensors in a bucket-aligned
    layout; anything else keeps the all-reduce path (and creates nothing)."""
    from multimodal_plankton_recognition_b200 import CLIPLoss
    mod = CLIPLoss(sharded=True)
    assert mod._peer_scalars(torch.randn(8, 4), 2) is None and mod._xgpu is None and not mod._xgpu_tried


def test_upload_slices_cover_every_query_once():
    """GpuExactIndex.search_host: equal slices of whole 256-query units, the last one ragged."""
    from multimodal_plankton_recognition_b200.ann import upload_slice_rows
    for nq, d, sb in ((100000, 512, 32 << 20), (1111, 128, 256 * 128 * 4 + 7), (256, 64, 1), (5, 8, 1 << 30),
                      (70000, 200, 10 << 20)):
        rows = upload_slice_rows(nq, nq * d * 4, sb)
        assert rows > 0 and rows % 256 == 0
        starts = list(range(0, nq, rows))
        assert sum(min(rows, nq - s) for s in starts) == nq
        # no more slices than asked for (rounding up to 256 rows only ever makes slices larger)
        assert len(starts) <= max(1, -(-(nq * d * 4) // sb))
    assert upload_slice_rows(100000, 100000 * 512 * 4, 32 << 20) == 14336   # the C4 query matrix: 7 slices
