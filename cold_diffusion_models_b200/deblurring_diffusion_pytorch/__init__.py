"""same exports as the reference package deblurring_diffusion_pytorch/__init__.py"""
from ..unet import Unet
from ..deblurring import GaussianDiffusion
from ..trainer import Trainer

__all__ = ['GaussianDiffusion', 'Unet', 'Trainer']
