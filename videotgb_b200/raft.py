"""RAFT optical-flow extractor: drop-in for the reference `RAFT`
(src/models/components/xraft.py:52-158) as the offline pipeline uses it (eval/utils/model.py:75-84):

    of_extractor = RAFT(); of_extractor.load_state_dict(ckpt)          # same keys: fnet.*, cnet.*, update_block.*
    flow = of_extractor(frames[:-1], frames[1:])                        # [T-1, 2, H, W], iters=20, test_mode=True

Inference only.  Activations live as NHWC pixel rows; every convolution is an im2col kernel + the tcgen05 GEMM (bf16
operands, fp32 accumulation, fp32 outputs), batch norms (cnet, eval statistics) are folded into the convolution
weights, the z and r gate convolutions of each GRU half share one GEMM, the all-pairs correlation volume is one GEMM
per frame pair (fmap1 pre-scaled by 1/sqrt(256), exact in bf16), and the 4-level 9x9 lookup, gates, instance norm and
convex upsampling are the kernels of csrc/flow.cu.  The recurrent state (`net`, flow) stays in fp32.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .modeling import _Holder, _KernelModule


class _Conv2d(nn.Module):
    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        kh, kw = (k, k) if isinstance(k, int) else k
        self.weight = nn.Parameter(torch.empty(cout, cin, kh, kw))
        self.bias = nn.Parameter(torch.zeros(cout))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")
        self.stride = (stride, stride) if isinstance(stride, int) else stride
        self.padding = (padding, padding) if isinstance(padding, int) else padding


class _BatchNorm(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.eps = 1e-5


def _encoder(out_dim: int, batch_norm: bool) -> nn.Module:
    """Parameter layout of BasicEncoder (raft_utils/extractor.py:116-190)."""
    e = _Holder()
    e.conv1 = _Conv2d(3, 64, 7, 2, 3)
    if batch_norm:
        e.norm1 = _BatchNorm(64)
    cin = 64
    for li, dim in ((1, 64), (2, 96), (3, 128)):
        blocks = []
        for bi in (0, 1):
            b = _Holder()
            stride = 2 if (bi == 0 and li > 1) else 1
            b.conv1 = _Conv2d(cin if bi == 0 else dim, dim, 3, stride, 1)
            b.conv2 = _Conv2d(dim, dim, 3, 1, 1)
            if batch_norm:
                b.norm1, b.norm2 = _BatchNorm(dim), _BatchNorm(dim)
            if stride != 1:
                if batch_norm:
                    b.norm3 = _BatchNorm(dim)
                    b.downsample = nn.Sequential(_Conv2d(cin, dim, 1, stride, 0), b.norm3)     # norm3 under two names
                else:
                    b.downsample = nn.Sequential(_Conv2d(cin, dim, 1, stride, 0))
            blocks.append(b)
        setattr(e, f"layer{li}", nn.Sequential(*blocks))
        cin = dim
    e.conv2 = _Conv2d(128, out_dim, 1)
    return e


class InputPadder:
    """Replicate-pads [..., H, W] images up to the next multiple of 8 on both sides (xraft.py:29-49): centred padding in
    'sintel' mode, bottom-only rows otherwise."""

    def __init__(self, dims, mode="sintel"):
        self.ht, self.wd = dims[-2:]
        rows, cols = -self.ht % 8, -self.wd % 8
        top = rows // 2 if mode == "sintel" else 0
        self._pad = [cols // 2, cols - cols // 2, top, rows - top]          # F.pad order: left, right, top, bottom

    def pad(self, input):
        return F.pad(input, self._pad, mode="replicate")

    def unpad(self, x):
        left, right, top, bottom = self._pad
        return x[..., top: x.shape[-2] - bottom, left: x.shape[-1] - right]


def dp_state_to_normal(state_dict):
    """raft_utils/utils.py:85-91: a `torch.nn.DataParallel` checkpoint -> plain keys.  As in the reference, ONLY keys
    that start with `module` survive, with every `module.` removed."""
    return {k.replace("module.", ""): v for k, v in state_dict.items() if k.startswith("module")}


class RAFT(_KernelModule):
    def __init__(self):
        super().__init__()
        self.hidden_dim = self.context_dim = 128
        self.corr_levels, self.corr_radius = 4, 4
        self.fnet = _encoder(256, batch_norm=False)
        self.cnet = _encoder(256, batch_norm=True)
        u = self.update_block = _Holder()
        u.encoder = _Holder()
        u.encoder.convc1 = _Conv2d(324, 256, 1)
        u.encoder.convc2 = _Conv2d(256, 192, 3, 1, 1)
        u.encoder.convf1 = _Conv2d(2, 128, 7, 1, 3)
        u.encoder.convf2 = _Conv2d(128, 64, 3, 1, 1)
        u.encoder.conv = _Conv2d(256, 126, 3, 1, 1)
        u.gru = _Holder()
        for g in "zrq":
            setattr(u.gru, f"conv{g}1", _Conv2d(384, 128, (1, 5), 1, (0, 2)))
            setattr(u.gru, f"conv{g}2", _Conv2d(384, 128, (5, 1), 1, (2, 0)))
        u.flow_head = _Holder()
        u.flow_head.conv1 = _Conv2d(128, 256, 3, 1, 1)
        u.flow_head.conv2 = _Conv2d(256, 2, 3, 1, 1)
        u.mask = nn.Sequential(_Conv2d(128, 256, 3, 1, 1), nn.ReLU(), _Conv2d(256, 576, 1))

    # ---- weight packing: [Cout(pad 8), (ky,kx,c)(pad 8)] bf16 + fp32 bias, batch norm / scales folded in
    @staticmethod
    def _pack_conv(conv: _Conv2d, bn: _BatchNorm = None, scale: float = 1.0, cin_pad: int = None, extra=None):
        w = conv.weight.detach().float()
        b = conv.bias.detach().float()
        if extra is not None:                                   # a second convolution on the same input: stack outputs
            w = torch.cat([w, extra.weight.detach().float()], 0)
            b = torch.cat([b, extra.bias.detach().float()], 0)
        if bn is not None:
            g = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            w = w * g.view(-1, 1, 1, 1)
            b = (b - bn.running_mean.detach().float()) * g + bn.bias.detach().float()
        w, b = w * scale, b * scale
        cout, cin, kh, kw = w.shape
        if cin_pad is not None and cin_pad > cin:               # input rows carry zero-padded channels
            w = F.pad(w, (0, 0, 0, 0, 0, cin_pad - cin))
            cin = cin_pad
        w2 = w.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)  # column (ky*kw + kx)*C + c
        kpad, npad = (w2.shape[1] + 7) // 8 * 8, (cout + 7) // 8 * 8
        w2 = F.pad(w2, (0, kpad - w2.shape[1], 0, npad - cout))
        b = F.pad(b, (0, npad - cout))
        return SimpleNamespace(w=w2.to(torch.bfloat16).contiguous(), b=b.contiguous(), cout=cout, cin=cin, kh=kh, kw=kw,
                               stride=conv.stride, padding=conv.padding)

    def _pack(self):
        if self._packed is not None:
            return self._packed
        pk = SimpleNamespace()
        for name, enc, bn in (("fnet", self.fnet, False), ("cnet", self.cnet, True)):
            e = SimpleNamespace(conv1=self._pack_conv(enc.conv1, enc.norm1 if bn else None), blocks=[],
                                conv2=self._pack_conv(enc.conv2))
            for li in (1, 2, 3):
                for blk in getattr(enc, f"layer{li}"):
                    b = SimpleNamespace(conv1=self._pack_conv(blk.conv1, blk.norm1 if bn else None),
                                        conv2=self._pack_conv(blk.conv2, blk.norm2 if bn else None), down=None)
                    if hasattr(blk, "downsample"):
                        b.down = self._pack_conv(blk.downsample[0], blk.norm3 if bn else None)
                    e.blocks.append(b)
            setattr(pk, name, e)
        u = self.update_block
        pk.convc1 = self._pack_conv(u.encoder.convc1, cin_pad=328)
        pk.convc2 = self._pack_conv(u.encoder.convc2)
        pk.convf1 = self._pack_conv(u.encoder.convf1)
        pk.convf2 = self._pack_conv(u.encoder.convf2)
        pk.conv = self._pack_conv(u.encoder.conv)
        pk.zr = [self._pack_conv(getattr(u.gru, f"convz{s}"), extra=getattr(u.gru, f"convr{s}")) for s in "12"]
        pk.q = [self._pack_conv(getattr(u.gru, f"convq{s}")) for s in "12"]
        pk.fh1 = self._pack_conv(u.flow_head.conv1)
        pk.fh2 = self._pack_conv(u.flow_head.conv2)
        pk.mask0 = self._pack_conv(u.mask[0])
        pk.mask2 = self._pack_conv(u.mask[2], scale=0.25)       # `mask = .25 * self.mask(net)` (update.py:143)
        self._packed = pk
        return pk

    # ---- building blocks ---------------------------------------------------------------------------------
    @staticmethod
    def _conv(x_bf16, n, h, w, c, pc, out=None, epi=ops.EPI_F32):
        """bf16 pixel rows [n*h*w, >=c] -> fp32 (or, with the fused-ReLU epilogue, bf16) [n*ho*wo, cout_pad]; -> (y, ho, wo)."""
        ho = (h + 2 * pc.padding[0] - pc.kh) // pc.stride[0] + 1
        wo = (w + 2 * pc.padding[1] - pc.kw) // pc.stride[1] + 1
        if pc.kh == 1 and pc.kw == 1 and pc.stride == (1, 1) and c % 8 == 0:
            a = x_bf16 if x_bf16.shape[1] == c else x_bf16[:, :c]
        else:
            a = ops.im2col_nhwc(x_bf16, n, h, w, c, pc.kh, pc.kw, pc.stride, pc.padding)
        return ops.gemm(a, pc.w, pc.b, epi, out=out), ho, wo

    def _encode(self, e, img_rows, n, h, w, instance: bool):
        """BasicEncoder on bf16 pixel rows [n*h*w, 8] (3 real channels) -> fp32 [n*(h/8)*(w/8), out]."""
        def norm_relu(y, hw, want_f32=True):
            if instance:
                return ops.instance_norm_nhwc(y, n, hw, True, want_f32, True)
            return ops.act(y, ops.ACT_RELU, want_f32=want_f32, want_bf16=True)
        y, h, w = self._conv(img_rows, n, h, w, 3, e.conv1)
        xf, xb = norm_relu(y, h * w)
        c = 64
        for b in e.blocks:
            y, ho, wo = self._conv(xb, n, h, w, c, b.conv1)
            _, yb = norm_relu(y, ho * wo, want_f32=False)
            y2, _, _ = self._conv(yb, n, ho, wo, b.conv1.cout, b.conv2)
            yf, _ = norm_relu(y2, ho * wo)
            if b.down is not None:
                d, _, _ = self._conv(xb, n, h, w, c, b.down)
                if instance:
                    xf, _ = ops.instance_norm_nhwc(d, n, ho * wo, False, True, False)
                else:
                    xf = d
            xf, xb = ops.act(xf, ops.ACT_RELU, y=yf, want_f32=True, want_bf16=True)
            h, w, c = ho, wo, b.conv1.cout
        out, _, _ = self._conv(xb, n, h, w, c, e.conv2)
        return out, h, w

    @torch.no_grad()
    def forward(self, image1, image2, iters=20, flow_init=None, upsample=True, test_mode=True):
        """Estimate optical flow between pairs of frames: fp32 [N,3,H,W] (0..255 scale) x2 -> flow_up fp32 [N,2,H,W]."""
        self._device_check(image1)
        if not test_mode:
            raise NotImplementedError("the B200 RAFT is inference only (test_mode=True returns the final flow_up)")
        if flow_init is not None:
            raise NotImplementedError("flow_init is not used on the VideoTGB path")
        n, _, H, W = image1.shape
        if H % 8 or W % 8:
            raise ValueError("image sides must be divisible by 8 (use InputPadder)")
        if ((H // 8) * (W // 8)) % 8 or H < 128 or W < 128:
            raise ValueError("the correlation GEMM needs (H/8)*(W/8) % 8 == 0 and the 4-level pyramid H, W >= 128")
        self._pack()
        image1, image2 = image1.float().contiguous(), image2.float().contiguous()
        if self._use_graph(n):                                   # ~800 launches per call: replay a captured CUDA graph
            return self._graphs.run(("raft", n, H, W, iters), [image1, image2],
                                    lambda a, b: (self._forward_impl(a, b, iters),))[0]
        return self._forward_impl(image1, image2, iters)

    def _forward_impl(self, image1, image2, iters):
        pk = self._pack()
        n, _, H, W = image1.shape
        dev = image1.device
        rows1 = ops.raft_prep_image(image1)
        rows2 = ops.raft_prep_image(image2)
        fmaps, h, w = self._encode(pk.fnet, torch.cat([rows1, rows2], 0), 2 * n, H, W, instance=True)
        hw = h * w
        # all-pairs correlation / sqrt(dim): one GEMM per pair (fmap1 pre-scaled by 1/16, exact in bf16)
        _, f1 = ops.act(fmaps[: n * hw], ops.ACT_NONE, scale=1.0 / 16.0, want_bf16=True)
        _, f2 = ops.act(fmaps[n * hw:], ops.ACT_NONE, want_bf16=True)
        corr = torch.empty((n * hw, hw), dtype=torch.float32, device=dev)
        for i in range(n):
            ops.gemm(f1[i * hw: (i + 1) * hw], f2[i * hw: (i + 1) * hw], None, ops.EPI_F32, out=corr[i * hw: (i + 1) * hw])
        pyr = [corr.view(n * hw, h, w)]
        for _ in range(self.corr_levels - 1):
            pyr.append(ops.corr_pool(pyr[-1]))
        cnet, _, _ = self._encode(pk.cnet, rows1, n, H, W, instance=False)
        m = n * hw
        net = torch.empty((m, 128), dtype=torch.float32, device=dev)
        # hx = [h | inp | motion(126) | flow(2)] as one bf16 buffer: the GRU convolutions read all 384 channels in place
        hx = torch.zeros((m, 384), dtype=torch.bfloat16, device=dev)
        ops.act(cnet[:, :128], ops.ACT_TANH, out_f32=net, out_bf16=hx[:, :128])
        ops.act(cnet[:, 128:256], ops.ACT_RELU, out_bf16=hx[:, 128:256])
        rhx = hx.clone()                                         # [r*h | inp | motion | flow] for the q convolution
        flow = torch.zeros((m, 2), dtype=torch.float32, device=dev)
        flow_b = torch.zeros((m, 8), dtype=torch.bfloat16, device=dev)
        corr_feat = torch.zeros((m, 328), dtype=torch.bfloat16, device=dev)
        cor_flo = torch.empty((m, 256), dtype=torch.bfloat16, device=dev)
        mask = None
        for _ in range(iters):
            ops.corr_lookup(pyr, flow, n, h, w, corr_feat)
            # BasicMotionEncoder (update.py:97-113); ReLU is fused into the GEMM epilogue, outputs land in their concat slots
            cb, _, _ = self._conv(corr_feat, n, h, w, 328, pk.convc1, epi=ops.EPI_BF16_RELU)
            self._conv(cb, n, h, w, 256, pk.convc2, out=cor_flo[:, :192], epi=ops.EPI_BF16_RELU)
            fb, _, _ = self._conv(flow_b, n, h, w, 2, pk.convf1, epi=ops.EPI_BF16_RELU)
            self._conv(fb, n, h, w, 128, pk.convf2, out=cor_flo[:, 192:256], epi=ops.EPI_BF16_RELU)
            self._conv(cor_flo, n, h, w, 256, pk.conv, out=hx[:, 256:384], epi=ops.EPI_BF16_RELU)   # 126 real channels + 2 pad
            hx[:, 382:384].copy_(flow_b[:, :2])
            rhx[:, 256:384].copy_(hx[:, 256:384])
            # SepConvGRU (update.py:31-59): horizontal then vertical half
            for s in (0, 1):
                zr, _, _ = self._conv(hx, n, h, w, 384, pk.zr[s])
                ops.gru_rh(zr, net, rhx[:, :128])
                q, _, _ = self._conv(rhx, n, h, w, 384, pk.q[s])
                ops.gru_update(zr, q, net, hx[:, :128])
            nb = hx[:, :128]
            t, _, _ = self._conv(nb, n, h, w, 128, pk.fh1, epi=ops.EPI_BF16_RELU)
            delta, _, _ = self._conv(t, n, h, w, 256, pk.fh2)
            ops.act(flow, ops.ACT_NONE, y=delta[:, :2], out_f32=flow, out_bf16=flow_b[:, :2])     # F(t+1) = F(t) + delta
        # the upsampling mask only matters for the last iteration (test_mode returns the final flow_up)
        t, _, _ = self._conv(hx[:, :128], n, h, w, 128, pk.mask0, epi=ops.EPI_BF16_RELU)
        mask, _, _ = self._conv(t, n, h, w, 256, pk.mask2)
        return ops.raft_upsample(mask, flow, n, h, w)
