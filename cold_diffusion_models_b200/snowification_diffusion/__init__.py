"""snowification/diffusion (== decolor-diffusion/diffusion) surface: GaussianDiffusion, Trainer, forward processes"""
from ..snowification import GaussianDiffusion, DeColorization, Snow, ForwardProcessBase
from ..trainer import SnowificationTrainer as Trainer, get_dataset
from ..unet import Unet
from .model import get_model

__all__ = ['GaussianDiffusion', 'Trainer', 'DeColorization', 'Snow', 'ForwardProcessBase', 'Unet', 'get_dataset', 'get_model']
