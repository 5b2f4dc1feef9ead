"""same exports as the reference package defading-generation-diffusion-pytorch/defading_diffusion_pytorch/__init__.py
(the reference reuses the package name `defading_diffusion_pytorch` for it; it is a different GaussianDiffusion)."""
from ..unet import Unet
from ..defading_generation import GaussianDiffusion
from ..trainer import DefadingGenerationTrainer as Trainer

__all__ = ['GaussianDiffusion', 'Unet', 'Trainer']
