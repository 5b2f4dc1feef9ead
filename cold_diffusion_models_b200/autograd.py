"""torch.autograd bridge: lets `loss.backward()` (Trainer.train, DB:1190-1196) drive the engine's
backward schedule.  The Function's gradient w.r.t. parameters is written by the kernels directly into the
flat gradient buffer whose views are the parameters' `.grad`, so nothing is returned to autograd for them."""
import torch


class UnetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, unet, x, time, *params):
        eng = unet.engine
        save = {}
        out = eng.forward(x, time, save=save)
        ctx.unet = unet
        ctx.save = save
        ctx.nparams = len(params)
        return out

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.unet.engine
        eng.attach_grads()
        eng.backward(ctx.save, dout)
        ctx.save = None
        return (None, None, None) + (None,) * ctx.nparams
