"""Drop-in PCNNet (reference: models/model.py:9-64) for the north-star pair enc_type='vn_pointnet' +
dec_type='vn_foldingnet', plus enc_type='vn_dgcnn_fps' (pairs with vn_foldingnet at latent_dim=512).  Same constructor (config namespace with num_coarse, latent_dim, only_coarse, device,
enc_pretrained), same forward(input, rot=None) -> (coarse, fine), same state_dict keys ('encoder.*', 'decoder.*')."""
from __future__ import annotations

import torch
import torch.nn as nn

from .dgcnn import VN_DGCNN_fps
from .pcn import Attention_VN_FoldingNet, VN_FoldingNet, VN_PointNet


def _make_vn_dgcnn_fps(config):
    return VN_DGCNN_fps(config, only_coarse=config.only_coarse)


# enc_type / dec_type strings of the reference's config.json -> builders of the sm_100a modules
_ENCODERS = {"vn_pointnet": VN_PointNet, "vn_dgcnn_fps": _make_vn_dgcnn_fps}
_DECODERS = {"vn_foldingnet": VN_FoldingNet, "attention_vn_foldingnet": Attention_VN_FoldingNet}


class PCNNet(nn.Module):
    """models/model.py:9-64: encoder -> (coarse, global feature) -> decoder -> fine, with the reference's optional frozen pretrained
    encoder (`enc_pretrained`), `only_coarse` mode and the 448-coarse output convention."""

    def __init__(self, config, enc_type="vn_pointnet", dec_type="vn_foldingnet"):
        super().__init__()
        if enc_type not in _ENCODERS:
            raise Exception(f"encoder type {enc_type} not supported yet (B200 path covers {sorted(_ENCODERS)}, SURVEY.md 8)")
        if not config.only_coarse and dec_type not in _DECODERS:
            raise Exception(f"decoder type {dec_type} not supported yet (B200 hot path covers {sorted(_DECODERS)}, SURVEY.md 8)")
        self.num_coarse = config.num_coarse
        self.only_coarse = config.only_coarse
        self.encoder = _ENCODERS[enc_type](config).to(config.device)
        if config.enc_pretrained != "none":      # frozen pretrained encoder (models/model.py:29-39)
            self.encoder.load_state_dict(torch.load(config.enc_pretrained), strict=False)
            self.encoder.requires_grad_(False)
        if not self.only_coarse:
            self.decoder = _DECODERS[dec_type](config).to(config.device)

    def forward(self, input, rot=None):
        coarse, feature_global = self.encoder(input)
        # num_coarse == 448: the encoder returns (224 predicted, 224 predicted + 224 sampled); the decoder folds the first, the loss sees the second
        seed_points, coarse_out = coarse if self.num_coarse == 448 else (coarse, coarse)
        if self.only_coarse:
            return coarse_out, None
        return coarse_out, self.decoder(seed_points, feature_global, rot)


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
