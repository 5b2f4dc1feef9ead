"""same exports as the reference package denoising_diffusion_pytorch/__init__.py"""
from ..unet import Unet
from ..denoising import GaussianDiffusion
from ..trainer import DenoisingTrainer as Trainer

__all__ = ['GaussianDiffusion', 'Unet', 'Trainer']
