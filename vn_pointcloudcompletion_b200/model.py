"""Drop-in PCNNet (reference: models/model.py:9-64) for the north-star pair enc_type='vn_pointnet' +
dec_type='vn_foldingnet', plus enc_type='vn_dgcnn_fps' (pairs with vn_foldingnet at latent_dim=512).  Same constructor (config namespace with num_coarse, latent_dim, only_coarse, device,
enc_pretrained), same forward(input, rot=None) -> (coarse, fine), same state_dict keys ('encoder.*', 'decoder.*')."""
from __future__ import annotations

import torch
import torch.nn as nn

from .dgcnn import VN_DGCNN_fps
from .pcn import Attention_VN_FoldingNet, VN_FoldingNet, VN_PointNet


class PCNNet(nn.Module):
    def __init__(self, config, enc_type="vn_pointnet", dec_type="vn_foldingnet"):
        super().__init__()
        self.num_coarse = config.num_coarse
        self.only_coarse = config.only_coarse
        if enc_type == "vn_pointnet":
            self.encoder = VN_PointNet(config).to(config.device)
        elif enc_type == "vn_dgcnn_fps":
            self.encoder = VN_DGCNN_fps(config, only_coarse=config.only_coarse).to(config.device)
        else:
            raise Exception(f"encoder type {enc_type} not supported yet (B200 path covers vn_pointnet and vn_dgcnn_fps, SURVEY.md 8)")
        if config.enc_pretrained != "none":
            sd = torch.load(config.enc_pretrained)
            self.encoder.load_state_dict(sd, strict=False)
            for param in self.encoder.parameters():
                param.requires_grad = False
        if not config.only_coarse:
            if dec_type == "vn_foldingnet":
                self.decoder = VN_FoldingNet(config).to(config.device)
            elif dec_type == "attention_vn_foldingnet":
                self.decoder = Attention_VN_FoldingNet(config).to(config.device)
            else:
                raise Exception(f"decoder type {dec_type} not supported yet (B200 hot path covers vn_foldingnet, SURVEY.md 8)")

    def forward(self, input, rot=None):
        coarse, feature_global = self.encoder(input)
        if self.num_coarse == 448:
            if self.only_coarse:
                return coarse[1], None
            fine = self.decoder(coarse[0], feature_global, rot)
            return coarse[1], fine
        if self.only_coarse:
            return coarse, None
        fine = self.decoder(coarse, feature_global, rot)
        return coarse, fine


class Rotate:
    """Minimal stand-in for pytorch3d.transforms.Rotate as the reference uses it (train.py:131-138, pcn.py:370):
    row-vector convention, transform_points(p) = p @ R with R [B,3,3]."""

    def __init__(self, R):
        self.R = R

    def transform_points(self, p):
        # p [.., n, 3] @ R [B, 3, 3] written as a broadcast multiply-add (a 16x3x3 product: no library GEMM on the path)
        return (p.unsqueeze(-1) * self.R.unsqueeze(-3)).sum(-2)


def random_rotations(B, device=None, generator=None):
    """uniform SO(3) via normalised Gaussian quaternions (pytorch3d.transforms.random_rotations semantics) -> [B,3,3]"""
    q = torch.randn(B, 4, device=device, generator=generator)
    q = q / q.norm(dim=1, keepdim=True)
    r, i, j, k = q.unbind(1)
    R = torch.stack([1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r),
                     2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r),
                     2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)], dim=1)
    return R.view(B, 3, 3)
