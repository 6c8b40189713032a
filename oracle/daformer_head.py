"""fp32 restatement of the consumer right after the hot path: MADM's ``DAFormerHead`` with the shipped decoder config
(``config_files/common/models/mtmadise_multi_lora.py:43-63``: MLP embeds 512->256, depthwise-separable ASPP fusion with
dilations (1, 6, 12, 18), BN + ReLU, 19 classes).  TEST INFRASTRUCTURE; PARITY UNPINNED (mmcv 1.3.7 absent).

Follows ``modeling/sem_seg_head/daformer_head.py``: ``MLP`` :401-411, ``ASPPModule`` / ``DepthwiseSeparableASPPModule``
:341-398, ``ASPPWrapper`` :414-479, ``DAFormerHead.forward`` :702-749, ``cls_seg`` :673-699.  mmcv semantics restated:
``ConvModule`` = conv (no bias when a norm follows) -> BN -> ReLU; ``DepthwiseSeparableConvModule`` = depthwise ConvModule
(3x3, groups=in) -> pointwise ConvModule (1x1), both with norm + act; ``resize`` = ``F.interpolate``.

The head is NOT part of the product ("unchanged head"); it is used to check the north-star gate that the
argmax segmentation computed from the product's features is >= 99.5 % pixel-identical to the one from the oracle's.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ConvModule(nn.Module):
    def __init__(self, cin, cout, k, padding=0, dilation=1, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding, dilation=dilation, groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class DepthwiseSeparableConvModule(nn.Module):
    def __init__(self, cin, cout, k, padding, dilation):
        super().__init__()
        self.depthwise_conv = ConvModule(cin, cin, k, padding=padding, dilation=dilation, groups=cin)
        self.pointwise_conv = ConvModule(cin, cout, 1)

    def forward(self, x):
        return self.pointwise_conv(self.depthwise_conv(x))


class ASPPWrapper(nn.Module):
    def __init__(self, in_channels, channels, dilations=(1, 6, 12, 18)):
        super().__init__()
        mods = []
        for d in dilations:  # sep=True: dilation 1 -> 1x1 ConvModule, others -> depthwise-separable 3x3 (:383-398)
            mods.append(ConvModule(in_channels, channels, 1) if d == 1 else
                        DepthwiseSeparableConvModule(in_channels, channels, 3, padding=d, dilation=d))
        self.aspp_modules = nn.ModuleList(mods)
        self.bottleneck = ConvModule(len(dilations) * channels, channels, 3, padding=1)

    def forward(self, x):
        return self.bottleneck(torch.cat([m(x) for m in self.aspp_modules], dim=1))  # :463-479 (pool=False, no context layer)


class MLP(nn.Module):
    def __init__(self, input_dim, embed_dim):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)

    def forward(self, x):
        return self.proj(x.flatten(2).transpose(1, 2).contiguous())  # :408-411


class DAFormerHead(nn.Module):
    def __init__(self, in_channels=(512, 512, 512, 512), in_keys=("s2", "s3", "s4", "s5"), channels=256, num_classes=19, embed_dims=256):
        super().__init__()
        self.in_keys = list(in_keys)
        self.embed_layers = nn.ModuleDict({str(i): MLP(c, embed_dims) for i, c in enumerate(in_channels)})
        self.fuse_layer = ASPPWrapper(embed_dims * len(in_channels), channels)
        self.conv_seg = nn.Conv2d(channels, num_classes, 1)
        nn.init.normal_(self.conv_seg.weight, std=0.01)  # init_cfg Normal(std=0.01) on conv_seg (:553)
        nn.init.zeros_(self.conv_seg.bias)

    def forward(self, input_dict):
        feats = input_dict["output_features"]
        x = [feats[k] for k in self.in_keys]  # transfer_input_dict_to_list :661-671
        n = x[-1].shape[0]
        os_size = x[0].shape[2:]
        cs = []
        for i, f in enumerate(x):  # :725-741
            c = self.embed_layers[str(i)](f).permute(0, 2, 1).contiguous().reshape(n, -1, f.shape[2], f.shape[3])
            if c.shape[2:] != os_size:
                c = F.interpolate(c, size=os_size, mode="bilinear", align_corners=False)
            cs.append(c)
        x = self.fuse_layer(torch.cat(cs, dim=1))  # :743
        return self.conv_seg(x)  # cls_seg :673-699 (dropout is identity in eval)


def build_head(seed: int = 4321, variant: str = "base") -> DAFormerHead:
    """Random-init head in eval mode with non-trivial BatchNorm running statistics and a classifier whose logits are O(1).
    variant 's0': in_channels[0] = 128, in_keys[0] = 's0' (mtmadise_cityscapes_rgb_to_depth_11.py:51-55): fuses on the 512^2 grid."""
    torch.manual_seed(seed)
    head = DAFormerHead() if variant == "base" else DAFormerHead(in_channels=(128, 512, 512, 512), in_keys=("s0", "s3", "s4", "s5"))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in head.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(1.0 + 0.2 * torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        head.conv_seg.weight.copy_(torch.randn(head.conv_seg.weight.shape, generator=g) * 0.2)
    head.eval()
    for p in head.parameters():
        p.requires_grad_(False)
    return head
