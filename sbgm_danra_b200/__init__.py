"""sbgm_danra_b200 -- B200-native (sm_100a) implementation of the SBGM_DANRA hot path.

Mirrors the reference's Python import surface (`sbgm.score_unet`, `sbgm.score_sampling`):

    from sbgm_danra_b200.score_unet import (ScoreNet, Encoder, Decoder, DecoderBlock, ImageSelfAttention,
        SinusoidalEmbedding, marginal_prob_std, diffusion_coeff, marginal_prob_std_fn, diffusion_coeff_fn, loss_fn)
    from sbgm_danra_b200.score_sampling import Euler_Maruyama_sampler, pc_sampler, ode_sampler, guided_score_fn

All arithmetic on the path runs in hand-written CUDA kernels reached through the C ABI declared in
`include/sbgm_b200.h`; torch supplies device memory, streams, CUDA graphs and torch.distributed.
"""
from ._lib import load_library  # noqa: F401

__all__ = ["load_library", "install_as_sbgm"]


def install_as_sbgm() -> None:
    """Make `import sbgm.score_unet` / `import sbgm.score_sampling` resolve to this package, so the
    reference's callers (training_main, generation_main, ...) run unmodified on the CUDA path."""
    import sys
    import types
    from . import score_sampling, score_unet
    pkg = sys.modules.get("sbgm")
    if pkg is None:
        pkg = types.ModuleType("sbgm")
        pkg.__path__ = []  # type: ignore[attr-defined]
        sys.modules["sbgm"] = pkg
    sys.modules["sbgm.score_unet"] = score_unet
    sys.modules["sbgm.score_sampling"] = score_sampling
    pkg.score_unet = score_unet            # type: ignore[attr-defined]
    pkg.score_sampling = score_sampling    # type: ignore[attr-defined]
