"""same exports as the reference package demixing_diffusion_pytorch/__init__.py"""
from ..unet import Unet
from ..demixing import GaussianDiffusion
from ..trainer import DemixingTrainer as Trainer

__all__ = ['GaussianDiffusion', 'Unet', 'Trainer']
