"""gic_b200: B200 (sm_100a) hot path of kawshik8/GAN-Image-Captioning behind the reference's own
Python interfaces (Generator/Decoder.sample, Discriminator.forward, get_losses, GANInstructor).

Host side mirrors the reference modules (same class names, signatures, state_dict keys and error
behaviour); all compute goes through the C ABI of libgic_b200.so (include/gic_b200.h).  There is no
CPU fallback: constructing the modules works anywhere, running them needs a B200.
"""
from . import _lib  # noqa: F401
from ._lib import GEMM_BF16, GEMM_FP32, GEMM_TF32, GicError  # noqa: F401

__all__ = ["GEMM_FP32", "GEMM_TF32", "GEMM_BF16", "GicError"]

_default_mode = GEMM_FP32


def set_gemm_mode(mode: int) -> None:
    """Precision of the dense contractions: GEMM_FP32 (exact, CUDA cores), GEMM_TF32 (tcgen05 kind::tf32, fp32
    accumulate) or GEMM_BF16 (as TF32, with bf16 operands on the discriminator's and the vocab-backward contractions)."""
    global _default_mode
    if mode not in (GEMM_FP32, GEMM_TF32, GEMM_BF16):
        raise ValueError("unknown GEMM mode %r" % (mode,))
    _default_mode = mode


def get_gemm_mode() -> int:
    return _default_mode
