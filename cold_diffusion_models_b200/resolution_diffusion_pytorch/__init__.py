"""same exports as the reference package resolution_diffusion_pytorch/__init__.py (the DDPM `Model` is a next-round row)"""
from ..unet import Unet
from ..resolution import GaussianDiffusion
from ..trainer import Trainer

__all__ = ['GaussianDiffusion', 'Unet', 'Trainer']
