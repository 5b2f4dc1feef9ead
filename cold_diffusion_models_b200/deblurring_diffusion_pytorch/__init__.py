"""same exports as the reference package deblurring_diffusion_pytorch/__init__.py"""
from ..unet import Unet
from ..deblurring import GaussianDiffusion
from ..trainer import Trainer
from ..model2 import Model

__all__ = ['GaussianDiffusion', 'Unet', 'Trainer', 'Model']
