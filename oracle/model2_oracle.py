"""ORACLE (test infrastructure only).  CPU fp32 functional restatement of the DDPM-style `Model`
(M2 = deblurring-diffusion-pytorch/deblurring_diffusion_pytorch/Model2.py:191-332) in eval mode (dropout inactive),
driven by a reference-format state_dict.  Pinned against the unmodified reference by tests/test_oracle_golden.py."""
import math
import torch
import torch.nn.functional as F


def swish(x):
    return x * torch.sigmoid(x)


def timestep_embedding(t, dim):
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    emb = t.float()[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + '.weight'], sd[p + '.bias'], eps=1e-6)


def resblock(sd, p, x, temb):
    h = F.conv2d(swish(gn(sd, p + '.norm1', x)), sd[p + '.conv1.weight'], sd[p + '.conv1.bias'], padding=1)
    h = h + F.linear(swish(temb), sd[p + '.temb_proj.weight'], sd[p + '.temb_proj.bias'])[:, :, None, None]
    h = F.conv2d(swish(gn(sd, p + '.norm2', h)), sd[p + '.conv2.weight'], sd[p + '.conv2.bias'], padding=1)
    if (p + '.nin_shortcut.weight') in sd:
        x = F.conv2d(x, sd[p + '.nin_shortcut.weight'], sd[p + '.nin_shortcut.bias'])
    elif (p + '.conv_shortcut.weight') in sd:
        x = F.conv2d(x, sd[p + '.conv_shortcut.weight'], sd[p + '.conv_shortcut.bias'], padding=1)
    return x + h


def attnblock(sd, p, x):
    h = gn(sd, p + '.norm', x)
    q = F.conv2d(h, sd[p + '.q.weight'], sd[p + '.q.bias'])
    k = F.conv2d(h, sd[p + '.k.weight'], sd[p + '.k.bias'])
    v = F.conv2d(h, sd[p + '.v.weight'], sd[p + '.v.bias'])
    b, c, hh, ww = q.shape
    w_ = torch.bmm(q.reshape(b, c, hh * ww).permute(0, 2, 1), k.reshape(b, c, hh * ww)) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2).permute(0, 2, 1)
    h = torch.bmm(v.reshape(b, c, hh * ww), w_).reshape(b, c, hh, ww)
    return x + F.conv2d(h, sd[p + '.proj_out.weight'], sd[p + '.proj_out.bias'])


def model_forward(sd, x, t, *, ch, num_resolutions, num_res_blocks):
    temb = timestep_embedding(t, ch)
    temb = F.linear(temb, sd['temb.dense.0.weight'], sd['temb.dense.0.bias'])
    temb = F.linear(swish(temb), sd['temb.dense.1.weight'], sd['temb.dense.1.bias'])
    hs = [F.conv2d(x, sd['conv_in.weight'], sd['conv_in.bias'], padding=1)]
    for lv in range(num_resolutions):
        for ib in range(num_res_blocks):
            h = resblock(sd, 'down.%d.block.%d' % (lv, ib), hs[-1], temb)
            if ('down.%d.attn.%d.norm.weight' % (lv, ib)) in sd:
                h = attnblock(sd, 'down.%d.attn.%d' % (lv, ib), h)
            hs.append(h)
        if lv != num_resolutions - 1:
            hp = F.pad(hs[-1], (0, 1, 0, 1), mode='constant', value=0)
            hs.append(F.conv2d(hp, sd['down.%d.downsample.conv.weight' % lv], sd['down.%d.downsample.conv.bias' % lv], stride=2))
    h = resblock(sd, 'mid.block_1', hs[-1], temb)
    h = attnblock(sd, 'mid.attn_1', h)
    h = resblock(sd, 'mid.block_2', h, temb)
    for lv in reversed(range(num_resolutions)):
        for ib in range(num_res_blocks + 1):
            h = resblock(sd, 'up.%d.block.%d' % (lv, ib), torch.cat([h, hs.pop()], dim=1), temb)
            if ('up.%d.attn.%d.norm.weight' % (lv, ib)) in sd:
                h = attnblock(sd, 'up.%d.attn.%d' % (lv, ib), h)
        if lv != 0:
            h = F.interpolate(h, scale_factor=2.0, mode='nearest')
            h = F.conv2d(h, sd['up.%d.upsample.conv.weight' % lv], sd['up.%d.upsample.conv.bias' % lv], padding=1)
    h = swish(gn(sd, 'norm_out', h))
    return F.conv2d(h, sd['conv_out.weight'], sd['conv_out.bias'], padding=1)
