"""A small stand-in for the reference's DepthUNet (RangeCLIP/src/depth_segmentation_model/model.py:59-117) with the same
interface for the loss path: ``forward(depth) -> (pixel_embeddings, temperature_text, temperature_image)``, the two
``log_temperature_*`` parameters (model.py:77-78) and a decoder that ends like the reference's
(utils/src/decoder.py:112-116: ``output_conv -> nearest interpolate to the target shape -> L2 normalise``).

It is NOT the ResNet-18-UNet+ASPP backbone (that stays the reference's PyTorch code, out of scope here): three
convolutions, enough to exercise the drop-in losses inside a real autograd / DDP / optimizer step with gradients flowing
into convolution weights.  ``skip_tail=True`` returns the ``output_conv`` result [B, D, H/2, W/2] for
``compute_loss_shared2x2`` instead of the upsampled, normalised tensor.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import rangeclip_b200 as R


class StandInDepthUNet(R.DepthCLIPLossMixin, nn.Module):
    def __init__(self, embedding_dim=512, width=32, temperature_text=0.07, temperature_image=0.1):
        super().__init__()
        self.enc1 = nn.Conv2d(1, width, 3, stride=2, padding=1)
        self.enc2 = nn.Conv2d(width, 2 * width, 3, padding=1)
        self.output_conv = nn.Conv2d(2 * width, embedding_dim, 3, padding=1)
        self.log_temperature_text = nn.Parameter(torch.log(torch.tensor(temperature_text)))
        self.log_temperature_image = nn.Parameter(torch.log(torch.tensor(temperature_image)))

    def forward(self, depth, skip_tail=False):
        H, W = depth.shape[-2:]
        x = F.relu(self.enc1(depth))
        x = F.relu(self.enc2(x))
        out = self.output_conv(x)                                   # [B, D, H/2, W/2]
        if not skip_tail:
            out = F.interpolate(out, size=(H, W), mode='nearest')   # decoder.py:113
            out = F.normalize(out, p=2, dim=1)                      # decoder.py:114
        return out, torch.exp(self.log_temperature_text), torch.exp(self.log_temperature_image)
