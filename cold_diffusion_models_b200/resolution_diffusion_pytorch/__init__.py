"""same exports as the reference package resolution_diffusion_pytorch/__init__.py"""
from ..unet import Unet
from ..resolution import GaussianDiffusion
from ..trainer import ResolutionTrainer as Trainer
from ..model2 import Model

__all__ = ['GaussianDiffusion', 'Unet', 'Trainer', 'Model']
