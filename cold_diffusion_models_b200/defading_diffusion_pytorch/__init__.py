"""same exports as the reference package defading_diffusion_pytorch/__init__.py"""
from ..unet import Unet
from ..defading import GaussianDiffusion
from ..trainer import DefadingTrainer as Trainer
from ..model2 import Model

__all__ = ['GaussianDiffusion', 'Unet', 'Trainer', 'Model']
