"""Oracle (CPU, fp32): functional restatements of the three networks on the DxMI sampler path.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Each function takes the reference's `state_dict` (plain dict of
tensors with the reference's key names) and mirrors one reference `forward`.
"""
import math

import torch
import torch.nn.functional as F


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _gn(sd, name, x, eps):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def _swish(x):
    return x * torch.sigmoid(x)


# ------------------------------------------------------------------------------------------------ DDPM U-Net


def ddpm_timestep_embedding(t, dim):
    """models/DxMI/unet_small.py:9-27 - [sin | cos], frequency step log(1e4)/(half-1)."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -step)
    ang = t.float()[:, None] * freqs[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)


def _ddpm_resblock(sd, p, x, temb, dropout_masks=None):
    """models/DxMI/unet_small.py:117-136.  Eval mode: dropout is the identity.  Training mode (:126-127, `self.dropout(h)` between
    swish(norm2) and conv2): the caller supplies the scaled keep mask per block (`dropout_masks[p]`, values 0 or 1/(1-p))."""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x, 1e-6)), padding=1)
    h = h + _lin(sd, p + ".temb_proj", _swish(temb))[:, :, None, None]
    h = _swish(_gn(sd, p + ".norm2", h, 1e-6))
    if dropout_masks is not None:
        h = h * dropout_masks[p]
    h = _conv(sd, p + ".conv2", h, padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x)
    return x + h


def _ddpm_attn(sd, p, x):
    """models/DxMI/unet_small.py:167-191 - single head, scale C^-0.5 applied after q.k, fp32 softmax."""
    hn = _gn(sd, p + ".norm", x, 1e-6)
    b, c, h, w = x.shape
    q = _conv(sd, p + ".q", hn).reshape(b, c, h * w)
    k = _conv(sd, p + ".k", hn).reshape(b, c, h * w)
    v = _conv(sd, p + ".v", hn).reshape(b, c, h * w)
    scores = torch.bmm(q.transpose(1, 2), k) * (int(c) ** (-0.5))  # [b, query, key]
    probs = torch.softmax(scores, dim=2)
    out = torch.bmm(v, probs.transpose(1, 2)).reshape(b, c, h, w)
    return x + _conv(sd, p + ".proj_out", out)


def ddpm_unet_forward(sd, x, t, *, ch=128, ch_mult=(1, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,),
                      resolution=32, dropout_masks=None):
    """models/DxMI/unet_small.py:292-332 (Model.forward)."""
    assert x.shape[2] == x.shape[3] == resolution
    temb = ddpm_timestep_embedding(t, ch)
    temb = _lin(sd, "temb.dense.1", _swish(_lin(sd, "temb.dense.0", temb)))
    n_levels = len(ch_mult)
    hs = [_conv(sd, "conv_in", x, padding=1)]
    res = resolution
    for lvl in range(n_levels):
        for blk in range(num_res_blocks):
            h = _ddpm_resblock(sd, f"down.{lvl}.block.{blk}", hs[-1], temb, dropout_masks)
            if res in attn_resolutions:
                h = _ddpm_attn(sd, f"down.{lvl}.attn.{blk}", h)
            hs.append(h)
        if lvl != n_levels - 1:
            # Downsample: zero-pad right/bottom by one, then 3x3 stride 2 (unet_small.py:69-73)
            hs.append(_conv(sd, f"down.{lvl}.downsample.conv", F.pad(hs[-1], (0, 1, 0, 1)), stride=2))
            res //= 2
    h = hs[-1]
    h = _ddpm_resblock(sd, "mid.block_1", h, temb, dropout_masks)
    h = _ddpm_attn(sd, "mid.attn_1", h)
    h = _ddpm_resblock(sd, "mid.block_2", h, temb, dropout_masks)
    for lvl in reversed(range(n_levels)):
        for blk in range(num_res_blocks + 1):
            h = _ddpm_resblock(sd, f"up.{lvl}.block.{blk}", torch.cat([h, hs.pop()], dim=1), temb, dropout_masks)
            if res in attn_resolutions:
                h = _ddpm_attn(sd, f"up.{lvl}.attn.{blk}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # unet_small.py:50-54
            h = _conv(sd, f"up.{lvl}.upsample.conv", h, padding=1)
            res *= 2
    return _conv(sd, "conv_out", _swish(_gn(sd, "norm_out", h, 1e-6)), padding=1)


# ------------------------------------------------------------------------------------------------ value / energy net


def igebm_forward(sd, x, prefix=""):
    """models/modules.py:142-163 (IGEBMEncoderV2.forward, keepdim=False, no spectral norm, no class embedding)
    with ResBlockV2.forward (modules.py:71-101).  Returns [B, 1]."""
    p = prefix
    out = F.leaky_relu(_conv(sd, p + "conv1", x, padding=1), 0.2)
    downsample = (True, False, True, False, True, False)
    for i in range(6):
        b = f"{p}blocks.{i}"
        h = F.leaky_relu(_conv(sd, b + ".conv1", out, padding=1), 0.2)
        h = _conv(sd, b + ".conv2", h, padding=1)
        skip = F.conv2d(out, sd[b + ".skip.0.weight"]) if (b + ".skip.0.weight") in sd else out
        h = h + skip
        if downsample[i]:
            h = F.avg_pool2d(h, 2)
        out = F.leaky_relu(h, 0.2)
    out = F.relu(out)
    out = out.view(out.shape[0], out.shape[1], -1).sum(2)
    out = _lin(sd, p + "linear", out)
    if (p + "out_scale.weight") in sd:
        out = _lin(sd, p + "out_scale", out)
    return out


def value_forward(sd, x, t=None):
    """models/value.py:8-12 - TimeIndependentValue drops t; keys carry the `net.` prefix."""
    return igebm_forward(sd, x, prefix="net.")


# ------------------------------------------------------------------------------------------------ ADM / EDM U-Net


def adm_timestep_embedding(t, dim, max_period=10000):
    """models/cm/nn.py:119-137 - [cos | sin], freqs = exp(-ln(max_period) * i / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn32(sd, name, x):
    """models/cm/nn.py:19-21 - GroupNorm32: compute in fp32 (eps 1e-5), cast back to the input dtype."""
    return F.group_norm(x.float(), 32, sd[name + ".weight"].float(), sd[name + ".bias"].float(), 1e-5).type(x.dtype)


def _adm_resblock(sd, p, x, emb, *, up=False, down=False, scale_shift=True):
    """models/cm/unet.py:240-260 (ResBlock._forward)."""
    h = F.silu(_gn32(sd, p + ".in_layers.0", x))
    if up:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif down:
        h = F.avg_pool2d(h, 2)
        x = F.avg_pool2d(x, 2)
    h = _conv(sd, p + ".in_layers.2", h, padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"]).type(h.dtype)
    emb_out = emb_out[:, :, None, None]
    if scale_shift:
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = _gn32(sd, p + ".out_layers.0", h) * (1 + scale) + shift
        h = _conv(sd, p + ".out_layers.3", F.silu(h), padding=1)
    else:
        h = h + emb_out
        h = _conv(sd, p + ".out_layers.3", F.silu(_gn32(sd, p + ".out_layers.0", h)), padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = _conv(sd, p + ".skip_connection", x)
    return x + h


def _adm_attention(sd, p, x, n_heads, half_softmax):
    """models/cm/unet.py:320-332 + QKVAttentionLegacy.forward :413-441 with the (three, heads, d) channel layout.
    `half_softmax` mirrors the reference's unconditional `.half()` (line 423): q, k, v, the logits and the softmax
    are evaluated in fp16.  With half_softmax=False the same math runs in the tensor's own dtype (fp32 oracle)."""
    b, c, hh, ww = x.shape
    qkv = F.conv2d(_gn32(sd, p + ".norm", x), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"]).view(b, 3 * c, hh * ww)
    if half_softmax:
        qkv = qkv.half()
    d = c // n_heads
    q, k, v = qkv.view(b, 3, n_heads, d, hh * ww).unbind(1)
    q = q.reshape(b * n_heads, d, -1)
    k = k.reshape(b * n_heads, d, -1)
    v = v.reshape(b * n_heads, d, -1)
    s = 1 / math.sqrt(math.sqrt(d))
    w = torch.einsum("bct,bcs->bts", q * s, k * s)
    w = torch.softmax(w, dim=-1).type(w.dtype)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, c, hh, ww)
    return x + F.conv2d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def adm_layout(*, image_size, model_channels, channel_mult, num_res_blocks, attention_ds, resblock_updown=True):
    """Block structure of UNetModel.__init__ (models/cm/unet.py:600-737): returns (input_blocks, output_blocks) as
    lists of lists of ('res'|'attn'|'down'|'up', in_ch, out_ch) in module-index order."""
    ch = int(channel_mult[0] * model_channels)
    inputs = [[("conv", None, ch)]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            out = int(mult * model_channels)
            layers = [("res", ch, out)]
            ch = out
            if ds in attention_ds:
                layers.append(("attn", ch, ch))
            inputs.append(layers)
            chans.append(ch)
        if level != len(channel_mult) - 1:
            inputs.append([("down", ch, ch)])
            chans.append(ch)
            ds *= 2
    outputs = []
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            out = int(model_channels * mult)
            layers = [("res", ch + ich, out)]
            ch = out
            if ds in attention_ds:
                layers.append(("attn", ch, ch))
            if level and i == num_res_blocks:
                layers.append(("up", ch, ch))
                ds //= 2
            outputs.append(layers)
    return inputs, outputs


def adm_unet_forward(sd, x, timesteps, y=None, *, image_size, model_channels, channel_mult, num_res_blocks,
                     attention_ds, num_head_channels=64, use_scale_shift_norm=True, fp16_torso=True, half_softmax=None):
    """models/cm/unet.py:761-790 (UNetModel.forward).  fp16_torso=True is the only mode the reference supports
    (convert_to_fp16 + QKVAttentionLegacy.half()): torso activations/conv weights fp16, GroupNorm and the
    embedding MLP fp32, head fp32.  fp16_torso=False evaluates the same graph entirely in fp32; `half_softmax` (default: follows
    fp16_torso) = True with fp16_torso=False is the reference built with use_fp16=False, whose attention still runs `.half()`."""
    if half_softmax is None:
        half_softmax = fp16_torso
    emb = adm_timestep_embedding(timesteps, model_channels)
    emb = F.linear(F.silu(_lin(sd, "time_embed.0", emb)), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if "label_emb.weight" in sd:
        assert y is not None and y.shape == (x.shape[0],)
        emb = emb + sd["label_emb.weight"][y]
    inputs, outputs = adm_layout(image_size=image_size, model_channels=model_channels, channel_mult=channel_mult,
                                 num_res_blocks=num_res_blocks, attention_ds=attention_ds)
    torso = torch.float16 if fp16_torso else torch.float32

    def run(prefix, layers, h):
        for j, (kind, cin, cout) in enumerate(layers):
            p = f"{prefix}.{j}"
            if kind == "conv":
                h = _conv(sd, p, h, padding=1)
            elif kind in ("res", "down", "up"):
                h = _adm_resblock(sd, p, h, emb, up=kind == "up", down=kind == "down", scale_shift=use_scale_shift_norm)
            else:
                h = _adm_attention(sd, p, h, cout // num_head_channels, half_softmax=half_softmax).type(h.dtype)
        return h

    h = x.type(torso)
    hs = []
    for i, layers in enumerate(inputs):
        h = run(f"input_blocks.{i}", layers, h)
        hs.append(h)
    mid_ch = h.shape[1]
    h = run("middle_block", [("res", mid_ch, mid_ch), ("attn", mid_ch, mid_ch), ("res", mid_ch, mid_ch)], h)
    for i, layers in enumerate(outputs):
        h = run(f"output_blocks.{i}", layers, torch.cat([h, hs.pop()], dim=1))
    h = h.type(x.dtype)
    return _conv(sd, "out.2", F.silu(_gn32(sd, "out.0", h)), padding=1)
