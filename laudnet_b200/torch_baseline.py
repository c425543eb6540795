"""Stock-PyTorch GPU baseline: the reference's masked-dense execution scheme on cuDNN, tuned the way a
PyTorch user would tune it - fp16 weights and activations cast ONCE, channels_last, cuDNN benchmark mode, the whole
forward captured in a CUDA graph.  This is a MEASUREMENT AID (bench.py's `gpu_baseline`, SURVEY.md 8d "the honest
GPU baseline"): nothing on the product path imports it, and it is not an oracle (fp16 arithmetic).

It restates the eval forward of LAUD-ResNet (reference laud_resnet.py:88-165, 316-363; maskers utils.py:47-65,
113-131) with stock torch ops only: every convolution runs densely and the result is multiplied by the 0/1 mask,
exactly as the reference does.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class TorchMaskedDenseResNet:
    def __init__(self, model, device, dtype=torch.float16):
        """`model`: a drop-in `laudnet_b200.ResNet` (used as a parameter container only)."""
        self.dev, self.dt = device, dtype
        cl = lambda w: w.detach().to(device=device, dtype=dtype).contiguous(memory_format=torch.channels_last)
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        bn = lambda m: tuple(t.detach().to(device=device, dtype=dtype) for t in (m.running_mean, m.running_var, m.weight, m.bias)) + (m.eps,)
        self.stem_w, self.stem_bn = cl(model.conv1.weight), bn(model.bn1)
        self.fc_w, self.fc_b = model.fc.weight.detach().to(device=device, dtype=dtype), model.fc.bias.detach().to(device=device, dtype=dtype)
        self.blocks = []
        for layer in (model.layer1, model.layer2, model.layer3, model.layer4):
            for blk in layer:
                d = dict(mode=blk.dyn_mode, stride=blk.stride, out_size=blk.output_size, mask_size=blk.mask_size,
                         G=blk.channel_dyn_group, width=blk.conv1.weight.shape[0],
                         w1=cl(blk.conv1.weight), w2=cl(blk.conv2.weight), w3=cl(blk.conv3.weight),
                         bn1=bn(blk.bn1), bn2=bn(blk.bn2), bn3=bn(blk.bn3), wd=None)
                if blk.downsample is not None:
                    d["wd"], d["bnd"] = cl(blk.downsample[0].weight), bn(blk.downsample[1])
                if blk.masker_channel is not None:
                    mk = blk.masker_channel
                    if not hasattr(mk, "layers"):
                        raise NotImplementedError("torch baseline: conv_linear masker")
                    if mk.layers == 2:
                        d["mk"] = (f32(mk.conv[0].weight), f32(mk.conv[0].bias), f32(mk.conv[2].weight), f32(mk.conv[2].bias))
                    else:
                        d["mk"] = (f32(mk.conv.weight), f32(mk.conv.bias))
                if blk.masker_spatial is not None:
                    d["ms"] = (f32(blk.masker_spatial.conv.weight), f32(blk.masker_spatial.conv.bias),
                               blk.spatial_mask_channel_group)
                self.blocks.append(d)

    @staticmethod
    def _bn(z, p):
        return F.batch_norm(z, p[0], p[1], p[2], p[3], False, 0.0, p[4])

    def forward(self, x):
        dt = self.dt
        z = F.conv2d(x, self.stem_w, stride=2, padding=3)
        x = F.max_pool2d(F.relu(self._bn(z, self.stem_bn)), 3, 2, 1)
        for d in self.blocks:
            b = x.shape[0]
            cm = None
            m3 = None
            if "mk" in d:                 # utils.py:113-131 (decision in fp32)
                pooled = x.float().mean(dim=(2, 3))
                mk = d["mk"]
                if len(mk) == 4:
                    logits = F.linear(F.relu(F.linear(pooled, mk[0], mk[1])), mk[2], mk[3])
                else:
                    logits = F.linear(pooled, mk[0], mk[1])
                G = d["G"]
                cm = (logits[:, :G] >= logits[:, G:]).to(dt).repeat_interleave(d["width"] // G, dim=1).view(b, -1, 1, 1)
            if "ms" in d:                 # utils.py:47-65 + laud_resnet.py:105-110
                w, bias, g = d["ms"]
                q = x.float()
                if d["mask_size"] < q.shape[2]:
                    q = F.adaptive_avg_pool2d(q, d["mask_size"])
                lg = F.conv2d(q, w, bias)
                small = (lg[:, :g] >= lg[:, g:]).to(dt)
                m3 = F.interpolate(small, size=d["out_size"], mode="nearest")
                if g > 1:
                    m3 = m3.repeat_interleave(d["w3"].shape[0] // g, dim=1)
            out = F.conv2d(x, d["w1"])
            if cm is not None:
                out = out * cm
            out = F.relu(self._bn(out, d["bn1"]))
            out = F.conv2d(out, d["w2"], stride=d["stride"], padding=1)
            if cm is not None:
                out = out * cm
            out = F.relu(self._bn(out, d["bn2"]))
            out = self._bn(F.conv2d(out, d["w3"]), d["bn3"])
            if m3 is not None:
                out = out * m3
            ident = x if d["wd"] is None else self._bn(F.conv2d(x, d["wd"], stride=d["stride"]), d["bnd"])
            x = F.relu(out + ident)
        feat = x.float().mean(dim=(2, 3)).to(dt)
        return F.linear(feat, self.fc_w, self.fc_b)

    def measure(self, x_nchw_f16: torch.Tensor, steps: int = 10, warmup: int = 3):
        """-> (images/s, ms per step, logits) of the CUDA-graphed forward at this batch."""
        prev = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = True
        try:
            xs = x_nchw_f16.to(self.dev, self.dt).contiguous(memory_format=torch.channels_last).clone()
            with torch.no_grad():
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        self.forward(xs)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    logits = self.forward(xs)
                for _ in range(warmup):
                    g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            return xs.shape[0] / ms * 1e3, ms, logits.float()
        finally:
            torch.backends.cudnn.benchmark = prev


class TorchMaskedDenseAdaViT:
    """Stock-PyTorch GPU baseline of BASELINE configs[3]: the AdaViT block executed MASKED-DENSE (everything computed, the
    0/1 decisions multiplied in - how a PyTorch implementation of AdaViT runs) with fp16 weights and activations,
    `F.scaled_dot_product_attention` with the key mask, captured in a CUDA graph.  A measurement aid (bench.py's
    `gpu_baseline` of --config 3); not an oracle (fp16 LayerNorm / policies change decisions)."""

    def __init__(self, model, device, dtype=torch.float16):
        self.dev, self.dt = device, dtype
        self.sd = {k: v.detach().to(device=device, dtype=dtype).contiguous() for k, v in model.state_dict().items()}
        self.D, self.H, self.depth, self.P = model.embed_dim, model.num_heads, model.depth, model.patch_size
        self.policy = [blk.has_policy for blk in model.blocks]
        self.flags = (model.ada_token, model.ada_head, model.ada_layer)

    def forward(self, img):
        s, D, H = self.sd, self.D, self.H
        d = D // H
        ln = lambda t, p: F.layer_norm(t, (D,), s[p + "weight"], s[p + "bias"], 1e-6)
        t = F.conv2d(img, s["patch_embed.proj.weight"], s["patch_embed.proj.bias"], stride=self.P).flatten(2).transpose(1, 2)
        x = torch.cat([s["cls_token"].expand(t.shape[0], -1, -1), t], 1) + s["pos_embed"]
        B, L, _ = x.shape
        for i in range(self.depth):
            p = f"blocks.{i}."
            tok = head = layer = None
            y = ln(x, p + "norm1.")
            if self.policy[i]:
                pt = ln(x[:, 0], p + "norm_policy.")
                if self.flags[2]:
                    layer = (F.linear(pt, s[p + "layer_select.weight"], s[p + "layer_select.bias"]) >= 0).to(x.dtype)
                if self.flags[1]:
                    head = (F.linear(pt, s[p + "head_select.weight"], s[p + "head_select.bias"]) >= 0).to(x.dtype)
                if self.flags[0]:
                    tl = F.linear(y[:, 1:], s[p + "token_select.weight"], s[p + "token_select.bias"]).squeeze(-1)
                    tok = torch.cat([torch.ones(B, 1, dtype=torch.bool, device=x.device), tl >= 0], 1)
            qkv = F.linear(y, s[p + "attn.qkv.weight"], s[p + "attn.qkv.bias"]).view(B, L, 3, H, d).permute(2, 0, 3, 1, 4)
            mask = tok[:, None, None, :] if tok is not None else None
            o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=mask)
            if head is not None:
                o = o * head[:, :, None, None]
            o = F.linear(o.transpose(1, 2).reshape(B, L, D), s[p + "attn.proj.weight"], s[p + "attn.proj.bias"])
            g = tok.to(x.dtype)[:, :, None] if tok is not None else 1.0
            x = x + o * g * (layer[:, 0, None, None] if layer is not None else 1.0)
            m = F.linear(F.gelu(F.linear(ln(x, p + "norm2."), s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"])),
                         s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"])
            x = x + m * g * (layer[:, 1, None, None] if layer is not None else 1.0)
        return F.linear(ln(x[:, 0], "norm."), s["head.weight"], s["head.bias"])

    def measure(self, x_nchw_f16: torch.Tensor, steps: int = 10, warmup: int = 3):
        """-> (img/s, ms per step, logits) of the CUDA-graphed forward."""
        with torch.no_grad():
            xs = x_nchw_f16.to(self.dt).clone()
            for _ in range(2):
                self.forward(xs)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.forward(xs)
            for _ in range(warmup):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
        return xs.shape[0] / (ms * 1e-3), ms, out.float()
