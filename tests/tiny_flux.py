"""A tiny 2-block MMDiT-shaped stand-in for diffusers' FluxTransformer2DModel (absent from this image), accepting the
exact keyword arguments the reference passes (SU:68-82, TR:134-144) and returning a 1-tuple.  BASELINE.json configs[0]:
"tiny random-init 2-block FluxTransformer, 256^2 latents, group size 4"."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _JointBlock(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.norm_i, self.norm_t = nn.LayerNorm(dim, elementwise_affine=False), nn.LayerNorm(dim, elementwise_affine=False)
        self.mod = nn.Linear(dim, 6 * dim)
        self.qkv_i, self.qkv_t = nn.Linear(dim, 3 * dim), nn.Linear(dim, 3 * dim)
        self.out_i, self.out_t = nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.mlp_i = nn.Sequential(nn.Linear(dim, 2 * dim), nn.GELU(), nn.Linear(2 * dim, dim))
        self.mlp_t = nn.Sequential(nn.Linear(dim, 2 * dim), nn.GELU(), nn.Linear(2 * dim, dim))

    def forward(self, img, txt, cond):
        sh1, sc1, g1, sh2, sc2, g2 = self.mod(F.silu(cond)).unsqueeze(1).chunk(6, dim=-1)
        B, Si, D = img.shape
        St = txt.shape[1]
        qi, ki, vi = self.qkv_i(self.norm_i(img) * (1 + sc1) + sh1).chunk(3, dim=-1)
        qt, kt, vt = self.qkv_t(self.norm_t(txt)).chunk(3, dim=-1)
        def heads(a):
            return a.view(B, -1, self.heads, D // self.heads).transpose(1, 2)
        q, k, v = heads(torch.cat([qt, qi], 1)), heads(torch.cat([kt, ki], 1)), heads(torch.cat([vt, vi], 1))
        att = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, St + Si, D)
        txt = txt + self.out_t(att[:, :St])
        img = img + g1 * self.out_i(att[:, St:])
        img = img + g2 * self.mlp_i(self.norm_i(img) * (1 + sc2) + sh2)
        txt = txt + self.mlp_t(self.norm_t(txt))
        return img, txt


class TinyFluxTransformer(nn.Module):
    def __init__(self, in_channels=64, dim=64, heads=4, depth=2, text_dim=32, pooled_dim=16):
        super().__init__()
        self.dim = dim
        self.x_in, self.t_in = nn.Linear(in_channels, dim), nn.Linear(text_dim, dim)
        self.time = nn.Sequential(nn.Linear(dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.guid = nn.Sequential(nn.Linear(dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.pool = nn.Linear(pooled_dim, dim)
        self.pos = nn.Linear(3, dim, bias=False)
        self.blocks = nn.ModuleList([_JointBlock(dim, heads) for _ in range(depth)])
        self.norm_out = nn.LayerNorm(dim, elementwise_affine=False)
        self.proj_out = nn.Linear(dim, in_channels)

    def _sin(self, t):
        half = self.dim // 2
        f = torch.exp(-math.log(10000.0) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
        a = t.float().view(-1, 1) * 1000.0 * f
        return torch.cat([a.sin(), a.cos()], dim=-1)

    def forward(self, hidden_states, encoder_hidden_states, timestep, guidance, txt_ids, pooled_projections, img_ids,
                joint_attention_kwargs=None, return_dict=False):
        B = hidden_states.shape[0]
        dt = self.x_in.weight.dtype
        cond = self.time(self._sin(timestep).to(dt)) + self.guid(self._sin(guidance.float().expand(B)).to(dt)) + self.pool(pooled_projections.to(dt))
        img = self.x_in(hidden_states.to(dt)) + self.pos(img_ids.to(dt)).unsqueeze(0)
        txt = self.t_in(encoder_hidden_states.to(dt)) + self.pos(txt_ids.to(dt)).unsqueeze(0)
        for blk in self.blocks:
            img, txt = blk(img, txt, cond)
        out = self.proj_out(self.norm_out(img))
        return (out,)
