"""snowification/diffusion/model: `get_model` and the two network families.  The reference's UnetConvNextBlock / UnetResNetBlock
have the same state_dict keys and the same forward as deblurring's Unet / Model (checked bit for bit on CPU), so they are those."""
from ...unet import Unet as UnetConvNextBlock
from ...model2 import Model as UnetResNetBlock
from .get_model import get_model

__all__ = ['get_model', 'UnetConvNextBlock', 'UnetResNetBlock']
