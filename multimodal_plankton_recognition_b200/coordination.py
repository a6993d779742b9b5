"""Drop-in for the hot-path class of reference src/coordination.py: ``CLIPLoss``.

Same constructor, parameter name/shape (``logit_scale``: 0-dim fp32, init 1.0, so reference
checkpoints load: key ``loss.logit_scale``), keyword call signature
``loss(image_emb=..., profile_emb=..., buckets=...)`` (reference src/model.py:95-98) and error
behaviour (AssertionError with the reference's message when the batch is not divisible by
``buckets``).  The arithmetic runs on the fused CUDA path (`ops.clip_loss`): the B x B logits are
never materialised, gradients flow to both raw embeddings and to ``logit_scale``.
"""
from __future__ import annotations

import os

import torch
from torch import Tensor
from torch.nn import Module, MSELoss, Parameter

from . import ops


class CLIPLoss(Module):
    """Symmetric CLIP-style InfoNCE (https://arxiv.org/abs/2103.00020), reference src/coordination.py:17-47.

    ``precision``: "bf16" (tcgen05 tensor-core path, <=2e-3 relative vs the reference at the
    reference's temperature), "fp16" (same kernels and speed with fp16 operands -- the reference's own
    '16-mixed' precision; 8x tighter, <=2e-3 at any temperature) or "fp32" (CUDA-core parity path,
    <=1e-5 relative).  Default: env ``PLK_PRECISION`` or "bf16".
    ``process_group``: if given (or ``sharded=True`` with the default group), the batch seen by
    ``forward`` is this rank's slice of a global batch and the loss is the global-batch loss
    (`dist.sharded_clip_loss`); the reference has no multi-GPU path, this is new functionality.
    """

    def __init__(self, bias: bool = False, *, precision: str | None = None, sharded: bool = False,
                 process_group=None) -> None:
        super().__init__()
        self.logit_scale = Parameter(torch.ones([]))
        precision = precision or os.environ.get("PLK_PRECISION", "bf16")
        if precision not in ops.MODES:
            raise ValueError(f"precision must be one of {sorted(ops.MODES)}, got {precision!r}")
        self.precision = precision
        self.sharded = sharded or process_group is not None
        self.process_group = process_group
        self._xgpu = None            # dist.XGpuScalars, created at the first bucket-aligned sharded forward
        self._xgpu_tried = False

    def forward(self, image_emb: Tensor, profile_emb: Tensor, buckets: int = 1) -> Tensor:
        # sharded: `buckets` counts the buckets of the GLOBAL batch; dist.sharded_fwd asserts on that
        assert self.sharded or image_emb.size(0) % buckets == 0, \
            "Batch size must be divisible by number of buckets!"
        if image_emb.dim() != 2 or image_emb.shape != profile_emb.shape:
            raise ValueError(f"expected two [B, d] embeddings of equal shape, got "
                             f"{tuple(image_emb.shape)} and {tuple(profile_emb.shape)}")
        mode = ops.MODES[self.precision]
        if self.sharded:
            from . import dist
            return dist.sharded_clip_loss(image_emb, profile_emb, self.logit_scale, int(buckets), mode,
                                          self.process_group, xgpu=self._peer_scalars(image_emb, int(buckets)))
        return ops.clip_loss(image_emb, profile_emb, self.logit_scale, int(buckets), mode)

    def forward_projected(self, image_feat: Tensor, profile_feat: Tensor, image_projection: Module,
                          profile_projection: Module, buckets: int = 1) -> Tensor:
        """the loss of ``image_projection(image_feat)`` / ``profile_projection(
        profile_feat)`` -- the two bias-free ``nn.Linear`` of reference src/model.py:29-30,:38-39 applied at
        :80-83 -- with each projection GEMM and the normalisation that opens the loss fused in one tcgen05
        kernel (`ops.project_normalise`).  Gradients flow to both feature tensors, both weights and
        ``logit_scale``.  Single GPU; `INTEGRATION.md` shows the call in ``training_step``."""
        for proj in (image_projection, profile_projection):
            if getattr(proj, "bias", None) is not None:
                raise ValueError("forward_projected fuses bias-free projections (the reference's nn.Linear(bias=False))")
        if image_feat.dim() != 2 or profile_feat.dim() != 2 or image_feat.size(0) != profile_feat.size(0):
            raise ValueError("expected [B, f_image] and [B, f_profile] features")
        assert image_feat.size(0) % buckets == 0, \
            "Batch size must be divisible by number of buckets!"
        if self.sharded:
            raise RuntimeError("forward_projected is single-GPU; project first and call the sharded loss")
        return ops.clip_loss_projected(image_feat, profile_feat, image_projection.weight, profile_projection.weight,
                                       self.logit_scale, int(buckets), ops.MODES[self.precision])

    def _peer_scalars(self, image_emb: Tensor, buckets: int):
        """The peer-memory scalar exchange for the bucket-aligned sharded case (every rank takes the same
        decision from the same shapes; creating it is a collective).  None -> NCCL all-reduces."""
        import torch.distributed as tdist
        if not image_emb.is_cuda or os.environ.get("PLK_XGPU", "1") == "0":
            return None
        world = tdist.get_world_size(self.process_group)
        n = image_emb.size(0)
        if world < 2 or world > 8 or (n * world) % buckets or n % ((n * world) // buckets):
            return None
        d = image_emb.size(1)
        if d % 128 or d > 1024:     # the fused exchange lives in the vectorised kernels (dist.xgpu_supported)
            return None
        if not self._xgpu_tried:
            self._xgpu_tried = True
            try:
                from .dist import XGpuScalars
                self._xgpu = XGpuScalars(image_emb.device, self.process_group)
            except Exception as e:      # no symmetric memory on this system
                import warnings
                warnings.warn(f"peer-memory scalar exchange unavailable ({e!r}); using NCCL all-reduces")
                self._xgpu = None
        return self._xgpu

    def graphed(self, image_emb: Tensor, profile_emb: Tensor, buckets: int = 1):
        """-> callable ``f(image_emb, profile_emb) -> loss`` whose forward AND backward replay CUDA graphs
        (`torch.cuda.make_graphed_callables`): ~40 us of host time per step instead of ~150.  The sample
        tensors fix shape and dtype; `buckets` is fixed too.  A sharded loss can be graphed in the bucket-aligned
        case (every bucket on one rank) when the peer-memory scalar exchange is available: the step then contains
        no collective launch -- the sums over the ranks run inside the kernels over NVLink.  Every rank must call
        this, and later the returned callable, the same number of times."""
        buckets = int(buckets)
        if self.sharded:
            if self._peer_scalars(image_emb, buckets) is None:
                raise RuntimeError("graphed() on a sharded loss needs the bucket-aligned case with the peer-memory "
                                   "scalar exchange (d % 128 == 0, NVLink symmetric memory): the general row-sharded "
                                   "step issues NCCL collectives, which are kept out of these captures")
        # a positional wrapper is graphed (make_graphed_callables rebinds the forward of the module it is
        # given and passes tensors positionally); `self` stays usable in eager mode and shares logit_scale
        return torch.cuda.make_graphed_callables(_Positional(self, buckets),
                                                 (image_emb.detach().clone().requires_grad_(),
                                                  profile_emb.detach().clone().requires_grad_()))

    def extra_repr(self) -> str:
        return f"precision={self.precision}, sharded={self.sharded}"


class _Positional(Module):
    def __init__(self, inner: Module, buckets: int = 1) -> None:
        super().__init__()
        self.inner = inner
        self.buckets = buckets

    def forward(self, image_emb: Tensor, profile_emb: Tensor) -> Tensor:
        return self.inner(image_emb=image_emb, profile_emb=profile_emb, buckets=self.buckets)


class SigLIPLoss(Module):
    """Pairwise-sigmoid coordination loss (https://arxiv.org/abs/2303.15343), drop-in for reference
    src/coordination.py:67-95: parameters ``logit_scale`` (init 1.0) and ``bias`` (init -10.0), both
    0-dim fp32; ``loss(image_emb=..., profile_emb=..., buckets=...)``.  Same similarity mainloop as
    `CLIPLoss` with a softplus epilogue -- no row / column statistics, so the forward is one pass and
    the recompute backward needs nothing from it.  Single GPU."""

    def __init__(self, *, precision: str | None = None) -> None:
        super().__init__()
        self.logit_scale = Parameter(torch.ones([]))
        self.bias = Parameter(-10 * torch.ones([]))
        precision = precision or os.environ.get("PLK_PRECISION", "bf16")
        if precision not in ops.MODES:
            raise ValueError(f"precision must be one of {sorted(ops.MODES)}, got {precision!r}")
        self.precision = precision

    def forward(self, image_emb: Tensor, profile_emb: Tensor, buckets: int = 1) -> Tensor:
        assert image_emb.size(0) % buckets == 0, \
            "Batch size must be divisible by number of buckets!"
        if image_emb.dim() != 2 or image_emb.shape != profile_emb.shape:
            raise ValueError(f"expected two [B, d] embeddings of equal shape, got "
                             f"{tuple(image_emb.shape)} and {tuple(profile_emb.shape)}")
        return ops.siglip_loss(image_emb, profile_emb, self.logit_scale, self.bias, int(buckets),
                               ops.MODES[self.precision])

    def extra_repr(self) -> str:
        return f"precision={self.precision}"


class SigLIPPlus(Module):
    """Drop-in for reference src/coordination.py:98-112: the fused SigLIP term plus ``beta`` times the MSE
    between the RAW embeddings; parameter paths ``siglip.logit_scale`` / ``siglip.bias``."""

    def __init__(self, beta: float = 0.25, *, precision: str | None = None) -> None:
        super().__init__()
        self.siglip = SigLIPLoss(precision=precision)
        self.l2 = MSELoss()
        self.beta = beta

    def forward(self, image_emb: Tensor, profile_emb: Tensor, buckets: int = 1) -> Tensor:
        return self.siglip(image_emb, profile_emb, buckets) + self.beta * self.l2(image_emb, profile_emb)


class CLIPPlus(Module):
    """Drop-in for reference src/coordination.py:50-64: the fused InfoNCE term
    plus ``beta`` times the MSE between the RAW embeddings.  The MSE term is one elementwise pass and
    stays in PyTorch; parameter path ``clip.logit_scale`` matches the reference's checkpoints."""

    def __init__(self, beta: float = 0.25, *, precision: str | None = None, sharded: bool = False,
                 process_group=None) -> None:
        super().__init__()
        self.clip = CLIPLoss(precision=precision, sharded=sharded, process_group=process_group)
        self.l2 = MSELoss()
        self.beta = beta

    def forward(self, image_emb: Tensor, profile_emb: Tensor, buckets: int = 1) -> Tensor:
        return self.clip(image_emb, profile_emb, buckets) + self.beta * self.l2(image_emb, profile_emb)
