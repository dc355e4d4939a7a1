"""TEST INFRASTRUCTURE ONLY -- plain-torch functional restatement of the reference RAFT optical-flow extractor
(SURVEY.md section 8f N1), state-dict in / tensor out, inference (eval) semantics.

  src/models/components/xraft.py:102-158                RAFT.forward (iters=20, test_mode=True -> flow_up)
  src/models/components/raft_utils/extractor.py:6-57    ResidualBlock, :116-190 BasicEncoder (fnet: instance norm, cnet: batch norm)
  src/models/components/raft_utils/corr.py:12-60        CorrBlock: all-pairs correlation / sqrt(dim), 4-level avg-pool pyramid,
                                                        9x9 bilinear lookup (x takes the FIRST window index, as the reference)
  src/models/components/raft_utils/update.py:9-144      FlowHead, SepConvGRU, BasicMotionEncoder, BasicUpdateBlock
  src/models/components/raft_utils/utils.py:62-90       bilinear_sampler, coords_grid
Pinned by tests/golden/raft.npz, produced by running the reference class (oracle/gen_golden.py raft).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _norm(sd, name, x, kind):
    if kind == "instance":
        return F.instance_norm(x, eps=1e-5)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=False, eps=1e-5)


def _res_block(sd, p, x, kind, stride):
    y = F.relu(_norm(sd, p + "norm1", _conv(sd, p + "conv1", x, stride, 1), kind))
    y = F.relu(_norm(sd, p + "norm2", _conv(sd, p + "conv2", y, 1, 1), kind))
    if stride != 1:
        x = _norm(sd, p + "downsample.1" if kind == "batch" else p + "norm3", _conv(sd, p + "downsample.0", x, stride, 0), kind)
    return F.relu(x + y)


def encoder(sd, prefix, x, kind):
    """BasicEncoder (extractor.py:116-190): [N,3,H,W] -> [N,out,H/8,W/8]."""
    x = F.relu(_norm(sd, prefix + "norm1", _conv(sd, prefix + "conv1", x, 2, 3), kind))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = _res_block(sd, f"{prefix}layer{li}.0.", x, kind, stride)
        x = _res_block(sd, f"{prefix}layer{li}.1.", x, kind, 1)
    return _conv(sd, prefix + "conv2", x)


def corr_pyramid(fmap1, fmap2, levels=4):
    b, d, h, w = fmap1.shape
    corr = torch.matmul(fmap1.view(b, d, h * w).transpose(1, 2), fmap2.view(b, d, h * w))
    corr = (corr / torch.sqrt(torch.tensor(d).float())).reshape(b * h * w, 1, h, w)
    pyr = [corr]
    for _ in range(levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def bilinear_sampler(img, coords):
    h, w = img.shape[-2:]
    xg, yg = coords.split([1, 1], dim=-1)
    grid = torch.cat([2 * xg / (w - 1) - 1, 2 * yg / (h - 1) - 1], dim=-1)
    return F.grid_sample(img, grid, align_corners=True)


def corr_lookup(pyr, coords, radius=4):
    coords = coords.permute(0, 2, 3, 1)
    b, h1, w1, _ = coords.shape
    out = []
    for i, corr in enumerate(pyr):
        d = torch.linspace(-radius, radius, 2 * radius + 1)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).to(coords.device)
        cl = coords.reshape(b * h1 * w1, 1, 1, 2) / 2 ** i + delta.view(1, 2 * radius + 1, 2 * radius + 1, 2)
        out.append(bilinear_sampler(corr, cl).view(b, h1, w1, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def update_block(sd, net, inp, corr, flow):
    p = "update_block."
    cor = F.relu(_conv(sd, p + "encoder.convc1", corr))
    cor = F.relu(_conv(sd, p + "encoder.convc2", cor, 1, 1))
    flo = F.relu(_conv(sd, p + "encoder.convf1", flow, 1, 3))
    flo = F.relu(_conv(sd, p + "encoder.convf2", flo, 1, 1))
    out = F.relu(_conv(sd, p + "encoder.conv", torch.cat([cor, flo], dim=1), 1, 1))
    x = torch.cat([inp, out, flow], dim=1)
    for suffix, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([net, x], dim=1)
        z = torch.sigmoid(_conv(sd, p + "gru.convz" + suffix, hx, 1, pad))
        r = torch.sigmoid(_conv(sd, p + "gru.convr" + suffix, hx, 1, pad))
        q = torch.tanh(_conv(sd, p + "gru.convq" + suffix, torch.cat([r * net, x], dim=1), 1, pad))
        net = (1 - z) * net + z * q
    delta = _conv(sd, p + "flow_head.conv2", F.relu(_conv(sd, p + "flow_head.conv1", net, 1, 1)), 1, 1)
    mask = 0.25 * _conv(sd, p + "mask.2", F.relu(_conv(sd, p + "mask.0", net, 1, 1)))
    return net, mask, delta


def upsample_flow(flow, mask):
    n, _, h, w = flow.shape
    mask = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, 8 * h, 8 * w)


def coords_grid(n, h, w, device):
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(n, 1, 1, 1).to(device)


def raft_forward(sd, image1, image2, iters=20, return_aux=False):
    """-> flow_up [N,2,H,W]  (xraft.py:102-158 with test_mode=True, flow_init=None)."""
    image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
    image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
    fmaps = encoder(sd, "fnet.", torch.cat([image1, image2], dim=0), "instance").float()
    fmap1, fmap2 = torch.split(fmaps, [image1.shape[0]] * 2, dim=0)
    pyr = corr_pyramid(fmap1, fmap2)
    cnet = encoder(sd, "cnet.", image1, "batch")
    net, inp = torch.split(cnet, [128, 128], dim=1)
    net, inp = torch.tanh(net), torch.relu(inp)
    n, _, h, w = image1.shape
    coords0 = coords_grid(n, h // 8, w // 8, image1.device)
    coords1 = coords0.clone()
    flow_up = None
    for _ in range(iters):
        corr = corr_lookup(pyr, coords1)
        net, mask, delta = update_block(sd, net, inp, corr, coords1 - coords0)
        coords1 = coords1 + delta
        flow_up = upsample_flow(coords1 - coords0, mask)
    if return_aux:
        return flow_up, dict(fmap1=fmap1, fmap2=fmap2, cnet=cnet, flow_lowres=coords1 - coords0, corr0=pyr[0])
    return flow_up
