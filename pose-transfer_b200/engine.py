"""Explicit forward/backward schedules of the generator and the discriminator on the ptk kernels.

This is the "autograd glue" of the product: instead of recording a torch graph, the fixed U-Net / PatchGAN
topology of the reference (models/networks.py:175-288,329-357) is executed as a hand-scheduled sequence of
kernel launches over pre-allocated NHWC workspaces, with the backward pass written out explicitly.  All
concatenations of the reference are replaced by channel-slice writes into shared buffers, every activation
is applied by the producer of a tensor, and the affine warp runs fused (utils/pose_transform.py:16-92).
"""
import contextlib
import os

import torch

from . import kernels as K
from .kernels import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, Slice


# Stream-level concurrency of independent branches (pose encoder, weight gradients, content loss).  bench.py switches it
# off for its per-kernel timing pass so that each kernel's CUDA-event duration is exclusive.
STREAMS = True


# floats of split-K scratch per stream (partial output tiles of the small bottleneck layers; 64 MB)
SPLITK_SCRATCH = 1 << 24


def ceil4(x):
    return (x + 3) // 4 * 4


def pad_in_channels(c):
    """Channel padding of a tensor that feeds a conv (activations, or output gradients feeding a dgrad): multiples
    of 32 = one 128-byte TMA / swizzle row of fp32, so that the stems (21 / 18 / 42 channels) and the dgrads of the
    narrow heads (3 / 1 channels) run on the tensor cores."""
    return (c + 31) // 32 * 32


def _weights_version(module):
    """Staleness key of a module's GEMM-layout weight packs: the manual counter bumped by whoever writes the weights through
    raw pointers (FlatAdam's kernel, ParamArena.bind, load_state_dict) plus torch's own in-place version counters, which see
    in-place edits of the parameters themselves (`with torch.no_grad(): p.copy_(..)`, `nn.init.*_(p)`, clipping, EMA) through
    any view of the storage.  Edits through `p.data` bypass torch's counter: follow them with ParamArena.bump()."""
    v = getattr(module, "_ptk_weights_version", None)
    if v is None:
        return None
    return (v, sum(p._version for p in module.parameters()))


class ConvLayer:
    """A Conv2d / ConvTranspose2d bound to its parameters.  Keeps GEMM-layout copies of the weight."""

    def __init__(self, weight, bias, transposed, k, stride, pad):
        self.weight, self.bias = weight, bias
        self.transposed, self.k, self.stride, self.pad = transposed, k, stride, pad
        if transposed:
            self.cin, self.cout = weight.shape[0], weight.shape[1]
        else:
            self.cout, self.cin = weight.shape[0], weight.shape[1]
        self.cin_pad, self.cout_pad = pad_in_channels(self.cin), ceil4(self.cout)
        self.dy_pad = pad_in_channels(self.cout)   # channel count of the gradient buffer that feeds this layer's dgrad
        self.taps = k * k
        self.w_fwd = None    # [taps][cin_pad][cout_pad]
        self.w_bwd = None    # [taps][dy_pad][cin_pad]  (weights of the dgrad gather-conv)
        self.w_dgrad_k = None  # [taps][cin_pad][dy_pad]: K-major tensor-core operand of the dgrad (== w_fwd unless narrow)
        self._derived = None   # GEMM-layout master weights: the transposed copy
        self._plans = {}       # wgrad geometry -> number of split-K partial gradients
        self.impl = {"auto": K.IMPL_AUTO, "simt": K.IMPL_SIMT, "tc": K.IMPL_AUTO}[os.environ.get("PTK_CONV_IMPL", "auto")]

    def _master(self):
        """(flat weight [tap][A][B_pad], flat grad, A, B_pad, taps) if the parameter lives in the arena in GEMM layout
        (models/pose_gan.ParamArena) and that layout matches this layer's operand paddings, else None."""
        m = getattr(self.weight, "_ptk_master", None)
        if m is None:
            return None
        flat, grad, A, Bp, taps = m
        if flat.device != self.weight.device or self.weight.data_ptr() != flat.data_ptr():
            return None
        if self.transposed:      # [tap][Cin][Cout_pad] == w_fwd ; derived w_bwd = [tap][dy_pad][cin_pad]
            ok = A == self.cin_pad and Bp == self.cout_pad and Bp == self.dy_pad
        else:                    # [tap][Cout][Cin_pad] == w_bwd ; derived w_fwd = [tap][cin_pad][cout_pad]
            ok = A == self.dy_pad and A == self.cout_pad and Bp == self.cin_pad
        return m if ok and taps == self.taps else None

    def _alloc(self):
        if self.w_fwd is None or self.w_fwd.device != self.weight.device:
            dev = self.weight.device
            # zero-initialised once: the pack kernels never write the channel-padding entries
            self.w_fwd = torch.zeros(self.taps * self.cin_pad * self.cout_pad, device=dev)
            self.w_bwd = torch.zeros(self.taps * self.cin_pad * self.dy_pad, device=dev)
            self.w_dgrad_k = self.w_fwd if self.dy_pad == self.cout_pad else torch.zeros(self.taps * self.cin_pad * self.dy_pad, device=dev)

    def pack_forward(self):
        self._pack_layouts()

    def _pack_layouts(self):
        """Both GEMM layouts are refreshed in one pass: w_fwd = [tap][cin][cout] is the CUDA-core fprop operand AND the
        K-major tensor-core operand of the dgrad; w_bwd = [tap][cout][cin] is the reverse."""
        m = self._master()
        if m is not None:
            # the arena storage IS one of the two layouts; the other one is its per-tap transpose
            flat, _, A, Bp, taps = m
            if self._derived is None or self._derived.device != flat.device:
                self._derived = torch.empty(taps * A * Bp, device=flat.device)
            K.transpose_weight(flat, self._derived, taps, A, Bp)
            if self.transposed:
                self.w_fwd, self.w_bwd = flat, self._derived
            else:
                self.w_bwd, self.w_fwd = flat, self._derived
            self.w_dgrad_k = self.w_fwd
            return
        if self._derived is not None:     # parameter storage moved out of the arena: back to private copies
            self._derived, self.w_fwd = None, None
        self._alloc()
        w = self.weight.detach()
        if self.transposed:   # torch layout [A=Cin][B=Cout][k][k]: w_fwd = [t][a][b], w_bwd = [t][b][a]
            K.pack_weight_dual(w, self.w_fwd, self.w_bwd, self.cin, self.cout, self.taps, self.cin_pad, self.cout_pad,
                               self.dy_pad, self.cin_pad)
            if self.w_dgrad_k is not self.w_fwd:
                K.pack_weight(w, self.w_dgrad_k, self.cin, self.cout, self.taps, self.cin_pad, self.dy_pad, 0)
        else:                 # torch layout [A=Cout][B=Cin][k][k]: w_bwd = [t][a][b], w_fwd = [t][b][a]
            K.pack_weight_dual(w, self.w_bwd, self.w_fwd, self.cout, self.cin, self.taps, self.dy_pad, self.cin_pad,
                               self.cin_pad, self.cout_pad)
            if self.w_dgrad_k is not self.w_fwd:
                K.pack_weight(w, self.w_dgrad_k, self.cout, self.cin, self.taps, self.dy_pad, self.cin_pad, 1)

    def pack_backward(self):
        """Kept for API symmetry: pack_forward() already refreshes every layout."""
        if self.w_bwd is None:
            self.pack_forward()

    def out_hw(self, H, W):
        if self.transposed:
            return (H - 1) * self.stride - 2 * self.pad + self.k, (W - 1) * self.stride - 2 * self.pad + self.k
        return (H + 2 * self.pad - self.k) // self.stride + 1, (W + 2 * self.pad - self.k) // self.stride + 1

    def forward(self, x, N, H, W, y, act=ACT_NONE, stats=None, y_nchw=None, scratch=None):
        """x: Slice with cin_pad readable channels; y: Slice (cout channels) or None.  scratch: fp32 buffer for the
        deterministic split-K of layers with few output tiles (one per stream that may run this layer)."""
        OH, OW = self.out_hw(H, W)
        g = K.conv_geom(N, H, W, self.cin_pad, x.ld, OH, OW, self.cout, y.ld if y is not None else self.cout, self.k,
                        self.stride, self.pad, self.transposed, self.impl)
        K.conv_forward(g, x, self.w_fwd, self.w_bwd, self.bias.detach() if self.bias is not None else None, act, y, y_nchw, stats,
                       scratch)
        return OH, OW

    def dgrad(self, dy, N, H, W, dx, dx_channels=None, scratch=None):
        """dy: Slice over the OUTPUT grid (dy_pad readable channels); dx: Slice over the input grid (H, W)."""
        OH, OW = self.out_hw(H, W)
        cout_dx = self.cin if dx_channels is None else dx_channels
        g = K.conv_geom(N, OH, OW, self.dy_pad, dy.ld, H, W, cout_dx, dx.ld, self.k, self.stride, self.pad,
                        not self.transposed, self.impl)
        # the dgrad weights are [taps][dy_pad][cin_pad]; the kernel's inner extent is ceil4(Cout')
        assert ceil4(cout_dx) == self.cin_pad
        K.conv_forward(g, dy, self.w_bwd, self.w_dgrad_k, None, ACT_NONE, dx, None, None, scratch)

    def wgrad(self, x, dy, N, H, W, scratch, grad_w, accumulate=False):
        """grad_w (torch layout, fp32 view into the gradient arena) += dW.  On the trainer's in-place path (grad_w is the
        zero-filled GEMM-layout arena gradient) the kernel OVERWRITES unless accumulate=True (stacked generator: the same
        weights receive one contribution per stack)."""
        OH, OW = self.out_hw(H, W)
        g = K.conv_geom(N, H, W, self.cin_pad, x.ld, OH, OW, self.cout if self.cout <= 4 else self.cout_pad, dy.ld, self.k,
                        self.stride, self.pad, self.transposed, self.impl)
        if self.transposed:
            A, B, B_pad = self.cin, self.cout, self.cout_pad
            a_rows = self.cin_pad
        else:
            A, B, B_pad = self.cout, self.cin, self.cin_pad
            a_rows = self.cout if self.cout <= 4 else self.cout_pad
        n = self.taps * a_rows * B_pad
        m = self._master()
        if m is not None and m[1].numel() == n and grad_w.data_ptr() == m[1].data_ptr():
            # Trainer path: grad_w IS the arena gradient of this weight (GEMM layout = the kernel's output layout), which
            # the trainer zero-filled at the start of the update, so the kernel writes it in place.  Any other grad_w
            # (the autograd Functions of models/networks.py pass fresh tensors) takes the accumulate path below.
            gflat = m[1]
            key = (N, H, W, x.ld, dy.ld, scratch.numel())
            nparts = self._plans.get(key)
            if nparts is None:
                nparts = self._plans[key] = K.conv_wgrad_plan(g, scratch.numel())
            if nparts == 1 and not accumulate:
                got = K.conv_wgrad_parts(g, x, dy, gflat)           # capacity == one gradient => written in place
                assert got == 1
            else:
                got = K.conv_wgrad_parts(g, x, dy, scratch)
                K.sum_parts(scratch, got, n, gflat, n, accumulate)
        elif a_rows == A:
            # split-K partial gradients land back to back in the scratch and are summed (fixed order) by the unpack
            assert grad_w.is_contiguous(), "wgrad: grad_w must be a contiguous torch-layout tensor (or the arena gradient)"
            nparts = K.conv_wgrad_parts(g, x, dy, scratch)
            K.unpack_weight_grad_parts(scratch, nparts, n, grad_w, A, B, self.taps, B_pad, True)
        else:  # padded row count (never hit for this network: every wide channel count is a multiple of 4)
            raise RuntimeError("wgrad: padded A rows unsupported")


class NormLayer:
    """Scalar-affine per-sample norm (InstanceNorm3d(1) on unsqueeze(1), models/networks.py:159,164-169)."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class _Workspace:
    """Named, lazily allocated device buffers that persist across steps (PyTorch owns the memory)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape, dtype=torch.float32, zero=False):
        t = self.bufs.get(name)
        shape = tuple(int(s) for s in shape)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self.bufs[name] = t
        elif zero:
            t.zero_()
        return t


class GeneratorEngine:
    """Deformable_Generator.forward / backward (models/networks.py:269-288 + encoder/decoder :193-250)."""

    def __init__(self, module):
        self.m = module
        # Encoder branches: (name, sub-module, first input channel, input channels, warped skips).  The deformable
        # generator has an appearance branch (warped) and a pose branch; the src_baseline Generator
        # (src_baseline/models/networks.py:238-254) is the same U-Net with ONE un-warped encoder over the whole input.
        if hasattr(module, "encoder_app"):
            P = module.pose_dim
            self.specs = [("app", module.encoder_app, 0, 3 + P, True), ("pose", module.encoder_pose, 3 + P, P, False)]
        else:
            self.specs = [("enc", module.encoder, 0, module.input_nc, False)]
        self.in_channels = sum(sp[3] for sp in self.specs)
        self.stage_modules = dict([("decoder", module.decoder)] + [(sp[0], sp[1]) for sp in self.specs])
        self.enc = list(module.nfilters_enc)
        self.dec = list(module.nfilters_dec)
        self.L = len(self.enc)
        self.image_size = tuple(module.image_size) if getattr(module, "image_size", None) is not None else None
        self.ws = None
        self.layers_built = False
        self.saved = None

    def deep_split(self):
        """First encoder level of the 'deep' gradient bucket (0 = do not split an encoder's bucket)."""
        return 4 if self.L >= 6 else 0

    def stage_parts(self, stage, part):
        """Sub-modules whose parameters form the gradient bucket (stage, part) announced by backward()'s on_stage."""
        mod = self.stage_modules[stage]
        if part is None:
            return [mod]
        nets = list(mod.net)
        sp = self.deep_split()
        return nets[sp:] if part == "deep" else nets[:sp]

    def fork(self):
        """A second execution context over the SAME layers and weight packs with a private workspace and saved state, so
        that several forward passes can be alive at once (stacked generator: one context per stack)."""
        ctx = _ForkedGeneratorEngine.__new__(_ForkedGeneratorEngine)
        ctx._parent, ctx.ws, ctx.saved = self, None, None
        return ctx

    # ------------------------------------------------------------------ parameter binding
    def _build_layers(self):
        m = self.m
        self.enc_conv = {}
        self.enc_norm = {}
        for name, e, _, _, _ in self.specs:
            convs, norms = [], []
            for i, mod in enumerate(e.net):
                if i == 0:
                    convs.append(ConvLayer(mod.weight, mod.bias, False, 3, 1, 1))
                    norms.append(None)
                else:
                    conv = mod.net[1]
                    convs.append(ConvLayer(conv.weight, None, False, 4, 2, 1))
                    norms.append(NormLayer(mod.net[2].weight, mod.net[2].bias) if len(mod.net) > 2 else None)
            self.enc_conv[name], self.enc_norm[name] = convs, norms
        self.dec_conv, self.dec_norm = [], []
        for j in range(self.L - 1):
            blk = m.decoder.net[j]
            self.dec_conv.append(ConvLayer(blk.net[1].weight, None, True, 4, 2, 1))
            self.dec_norm.append(NormLayer(blk.net[3].weight, blk.net[3].bias))
        fin = m.decoder.net[self.L]
        self.final_conv = ConvLayer(fin.weight, fin.bias, False, 3, 1, 1)
        self.all_convs = sum((self.enc_conv[sp[0]] for sp in self.specs), []) + self.dec_conv + [self.final_conv]
        self.layers_built = True

    def _side_stream(self, device):
        """Second CUDA stream for the pose encoder (PTK_STREAMS=0 or engine.STREAMS = False disables)."""
        if device.type != "cuda" or not STREAMS or os.environ.get("PTK_STREAMS", "1") == "0":
            return None
        st = getattr(self, "_side", None)
        if st is None or st.device != device:
            st = self._side = torch.cuda.Stream(device=device)
        return st

    def _ensure(self, device):
        if not self.layers_built or self.final_conv.weight is not self.m.decoder.net[self.L].weight:
            self._build_layers()
        if self.ws is None or self.ws.device != device:
            self.ws = _Workspace(device)

    def _fast_head(self):
        """The narrow output head (Conv2d(C -> 3, k3, p1) + tanh) runs as 1x1 tensor-core GEMMs over a 32-column
        'tap x channel' tensor (csrc/head.cu) unless the exact-fp32 CUDA-core mode is selected."""
        fc = self.final_conv
        return fc.impl != K.IMPL_SIMT and fc.cout <= 3 and fc.k == 3 and fc.cin % 32 == 0

    def pack_weights(self, backward=False):
        for c in self.all_convs:
            if c is self.final_conv and self._fast_head():
                continue
            c.pack_forward()
        if self._fast_head():
            fc = self.final_conv
            dev = fc.weight.device
            if getattr(self, "head_wk", None) is None or self.head_wk.device != dev:
                self.head_wk = torch.zeros(32 * fc.cin, device=dev)     # [32][Cin]: forward B operand
                self.head_wd = torch.zeros(fc.cin * 32, device=dev)     # [Cin][32]: dgrad B operand
            K.head_pack_weights(fc.weight.detach().contiguous(), self.head_wk, self.head_wd)
        self.packed_version = _weights_version(self.m)

    def pack_weights_backward(self):
        for c in self.all_convs:
            c.pack_backward()

    def _need_repack(self, repack):
        """repack=True: always; None: only if the weights moved since the last pack (_weights_version)."""
        if repack is None:
            v = _weights_version(self.m)
            return v is None or v != getattr(self, "packed_version", -1) or self.all_convs[0].w_fwd is None
        return bool(repack)

    # ------------------------------------------------------------------ geometry helpers
    def _sizes(self, H, W):
        hs, wsz = [H], [W]
        for _ in range(1, self.L):
            hs.append((hs[-1] + 2 - 4) // 2 + 1)
            wsz.append((wsz[-1] + 2 - 4) // 2 + 1)
        return hs, wsz

    def _cat_layout(self, j):
        """Channel layout of decoder level j's input: (dec_prev, one slot of enc_i channels per encoder branch), i = L-1-j."""
        i = self.L - 1 - j
        dprev = 0 if j == 0 else self.dec[j - 1]
        return i, dprev, dprev + len(self.specs) * self.enc[i]

    # ------------------------------------------------------------------ forward
    def forward(self, inp, warps, masks, drop=None, repack=True, d_input=None):
        """inp [N,3+2P,H,W] f32, warps [N,K,8] f32, masks [N,K,H,W] f64 -> out_gen [N,3,H,W] (fresh tensor).
        `inp` may also be a list of (tensor, first channel, channels) pieces whose concatenation along C is the input
        (the stacked generator feeds [previous output | pose_{i-1} | pose_i] without materialising the cat,
        models/networks.py:315-323).
        `drop`: list of three [N,512] (or [N,512,1,1]) Dropout2d noise tensors or None (=> drawn here).
        `d_input`: optional Slice of a discriminator input buffer that also receives out_gen (NHWC)."""
        pieces = [(inp, 0, inp.shape[1])] if torch.is_tensor(inp) else list(inp)
        pieces = [(t.contiguous(), c0, C) for t, c0, C in pieces]
        inp = pieces[0][0]
        assert all(t.dtype == torch.float32 and t.dim() == 4 for t, _, _ in pieces)   # device is enforced by the kernel wrappers
        N, _, H, W = inp.shape
        Ct = sum(C for _, _, C in pieces)
        assert N >= 2, "the reference's .squeeze() (models/networks.py:169) makes N=1 unsupported"
        L = self.L
        assert Ct == self.in_channels
        self._ensure(inp.device)
        ws = self.ws
        any_warp = any(sp[4] for sp in self.specs)
        if any_warp:
            warps = warps.contiguous().float()
            masks = masks.contiguous()
            if masks.dtype != torch.float64:
                masks = masks.double()
            Kp = warps.shape[1]
            H0, W0 = self.image_size
        else:
            Kp, H0, W0 = 0, H, W
        if self._need_repack(repack):
            self.pack_weights()
        hs, wsz = self._sizes(H, W)
        tag = "%d_%d_%d" % (N, H, W)
        n_norm = len(self.specs) * (L - 2) + (L - 1)
        stats = ws.get("stats" + tag, (n_norm, N, 2), torch.float64, zero=True)
        st_idx = {}

        def stat_for(key):
            if key not in st_idx:
                st_idx[key] = len(st_idx)
            return stats[st_idx[key]]

        # dropout noise (Dropout2d(0.5), models/networks.py:161; active in every forward of the reference, which never
        # calls .eval() -- a caller that does gets nn.Dropout2d's eval behaviour: identity)
        drops = []
        for j in range(min(3, L - 1)):
            if drop is None and not self.m.training:
                d = torch.ones(N, self.dec[j], 1, 1, device=inp.device)
            elif drop is None:
                d = torch.empty(N, self.dec[j], 1, 1, device=inp.device).bernoulli_(0.5).div_(0.5)
            else:
                d = drop[j].to(inp.device, torch.float32)
            drops.append(d.reshape(N, self.dec[j]).contiguous())

        # concat buffers (decoder inputs), everything stored post-ReLU
        cats = []
        for j in range(L):
            i, dprev, width = self._cat_layout(j)
            cats.append(ws.get("cat%d_%s" % (j, tag), (N, hs[i], wsz[i], width)))

        # mask pyramid for the 4 warped levels
        mlv = [ws.get("mask%d_%s" % (i, tag), (N, hs[i], wsz[i], Kp)) for i in range(min(4, L) if any_warp else 0)]
        if mlv:
            K.mask_pyramid_levels(masks, mlv)

        sv = {"N": N, "H": H, "W": W, "hs": hs, "ws": wsz, "tag": tag, "stats": stats, "st_idx": st_idx, "drops": drops,
              "cats": cats, "mlv": mlv, "warps": warps, "K": Kp, "z": {}, "act": {}, "yraw": {}, "argk": {}, "xin": {}}

        # The two encoders are independent until the decoder: the pose encoder runs on a side stream so that its
        # latency-bound bottleneck layers (16x16 .. 4x4, a handful of CTAs each) overlap the appearance encoder's.
        side = self._side_stream(inp.device)
        main = torch.cuda.current_stream() if side is not None else None
        if side is not None:
            side.wait_stream(main)
        for e_idx, (name, _, c_src0, cin, warped_branch) in enumerate(self.specs):
            with (torch.cuda.stream(side) if (side is not None and e_idx == 1) else contextlib.nullcontext()):
                convs, norms = self.enc_conv[name], self.enc_norm[name]
                ck = ws.get("splitk_side" if (side is not None and e_idx == 1) else "splitk_main", (SPLITK_SCRATCH,))
                warp_levels = []
                xin = ws.get("xin_%s_%s" % (name, tag), (N, H, W, convs[0].cin_pad))
                lo, segs = 0, []
                for t, t0, tc in pieces:        # channel range [c_src0, c_src0 + cin) of the (virtual) concatenation
                    a, b = max(lo, c_src0), min(lo + tc, c_src0 + cin)
                    if a < b:
                        segs.append((t, t0 + a - lo, b - a, a - c_src0))
                    lo += tc
                for q in range(0, len(segs), 4):   # whole rows incl. the zero channel padding, four source ranges per launch
                    if q == 0:
                        K.gather_nhwc(segs[:4], Slice(xin, 0, xin.shape[-1]), xin.shape[-1])
                    else:
                        for t, ts, tcn, td in segs[q:q + 4]:
                            K.nchw_to_nhwc(t, ts, tcn, Slice(xin, td, tcn))
                sv["xin"][name] = xin
                for i in range(L):
                    j = L - 1 - i
                    _, dprev, _ = self._cat_layout(j)
                    c = self.enc[i]
                    cat_slice = Slice(cats[j], dprev + e_idx * c, c)
                    z = ws.get("z_%s%d_%s" % (name, i, tag), (N, hs[i], wsz[i], c))
                    src = Slice(xin) if i == 0 else Slice(sv["act"][(name, i - 1)])
                    norm = norms[i]
                    st = stat_for((name, i)) if norm is not None else None
                    convs[i].forward(src, N, hs[i - 1] if i else H, wsz[i - 1] if i else W, Slice(z), ACT_NONE, st, scratch=ck)
                    sv["z"][(name, i)] = z
                    HW = hs[i] * wsz[i]
                    gam = norm.weight.detach() if norm is not None else None
                    bet = norm.bias.detach() if norm is not None else None
                    warped = warped_branch and i < 4
                    if i < L - 1:
                        act = ws.get("act_%s%d_%s" % (name, i, tag), (N, hs[i], wsz[i], c))
                        sv["act"][(name, i)] = act
                        if warped:
                            if norm is not None:
                                yraw = ws.get("yraw_%s%d_%s" % (name, i, tag), (N, hs[i], wsz[i], c))
                                K.gn_apply(z, st, gam, bet, None, N, HW, c, act, ACT_LEAKY, yraw, ACT_NONE)
                            else:
                                yraw = z
                                K.gn_apply(z, None, None, None, None, N, HW, c, act, ACT_LEAKY)
                            sv["yraw"][i] = yraw
                        else:
                            K.gn_apply(z, st, gam, bet, None, N, HW, c, act, ACT_LEAKY, cat_slice, ACT_RELU)
                    else:
                        K.gn_apply(z, st, gam, bet, None, N, HW, c, cat_slice, ACT_RELU)
                    if warped:
                        argk = ws.get("argk%d_%s" % (i, tag), (N, hs[i], wsz[i], c), torch.uint8)
                        sv["argk"][i] = argk
                        warp_levels.append(dict(x=Slice(sv["yraw"][i]), mask=mlv[i], y=cat_slice, argk=argk, C=c, h=hs[i], w=wsz[i]))
                        if i == min(4, L) - 1:
                            # all warped skip levels of this forward in ONE launch (utils/pose_transform.py:16-92 x 4)
                            K.warp_forward_levels(warp_levels, warps, N, Kp, H0, W0, ACT_RELU)
        if side is not None:
            main.wait_stream(side)

        # decoder
        zd = []
        for j in range(L - 1):
            i, dprev, width = self._cat_layout(j)
            co = self.dec[j]
            oh, ow = hs[i - 1], wsz[i - 1]
            z = ws.get("zd%d_%s" % (j, tag), (N, oh, ow, co))
            st = stat_for(("dec", j))
            self.dec_conv[j].forward(Slice(cats[j]), N, hs[i], wsz[i], Slice(z), ACT_NONE, st, scratch=ws.get("splitk_main", (SPLITK_SCRATCH,)))
            nl = self.dec_norm[j]
            K.gn_apply(z, st, nl.weight.detach(), nl.bias.detach(), drops[j] if j < 3 else None, N, oh * ow, co,
                       Slice(cats[j + 1], 0, co), ACT_RELU)
            zd.append(z)
        sv["zd"] = zd
        out = torch.empty(N, 3, H, W, device=inp.device)
        fc = self.final_conv
        if self._fast_head():
            z27 = ws.get("head_z" + tag, (N, H, W, 32))
            g1 = K.conv_geom(N, H, W, fc.cin, cats[L - 1].shape[-1], H, W, 32, 32, 1, 1, 0, False, K.IMPL_TC)
            K.conv_forward(g1, Slice(cats[L - 1]), None, self.head_wk, None, ACT_NONE, Slice(z27), None, None)
            K.head_shift_add(z27, fc.bias.detach(), fc.cout, ACT_TANH, out, d_input)
        else:
            fc.forward(Slice(cats[L - 1]), N, H, W, d_input, ACT_TANH, None, out)
        sv["out"] = out
        self.saved = sv
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, grads, dout_nchw=None, dout_nhwc=None, on_stage=None, accumulate=False, need_image_grad=False):
        """Accumulate parameter gradients into `grads` (dict: parameter -> fp32 tensor of the same shape).
        Uses the weight packs of the preceding forward().
        dout_nchw [N,3,H,W] and/or dout_nhwc (Slice over [N,H,W,*]) are gradients w.r.t. out_gen.
        on_stage(name) is called once every gradient of a sub-network ("decoder", "app", "pose") has been enqueued
        (the data-parallel trainer starts that bucket's all-reduce there).
        accumulate: add to the arena gradients instead of overwriting them (second and later stacks).
        need_image_grad: also return d loss / d input[:, 0:3] as NCHW [N,3,H,W] (the previous stack's output feeds these
        channels, models/networks.py:320-323); None otherwise."""
        sv = self.saved
        assert sv is not None, "forward() must run before backward()"
        ws, L = self.ws, self.L
        N, H, W, hs, wsz, tag = sv["N"], sv["H"], sv["W"], sv["hs"], sv["ws"], sv["tag"]
        cats, stats, st_idx, drops = sv["cats"], sv["stats"], sv["st_idx"], sv["drops"]
        H0, W0 = self.image_size if self.image_size is not None else (H, W)
        max_w = max(c.taps * c.cin_pad * c.cout_pad for c in self.all_convs)
        scratch = ws.get("wgrad_scratch", (max(4 * max_w, 1 << 24),))   # room for split-K partial gradients
        sums = ws.get("sums" + tag, tuple(stats.shape), torch.float64, zero=True)

        # final conv: tanh' then wgrad / bias grad / dgrad
        fc = self.final_conv
        dz4 = ws.get("dzf4" + tag, (N, H, W, 4))             # compact gradient w.r.t. the pre-tanh output (3 channels + pad)
        K.tanh_bwd_combine(dout_nchw, dout_nhwc, sv["out"], dz4, 4, N, 3, H, W)
        dcat = ws.get("dcat%d_%s" % (L - 1, tag), tuple(cats[L - 1].shape))
        if self._fast_head():
            dzs = ws.get("head_dzs" + tag, (N, H, W, 32))
            K.head_shift_gather(dz4, fc.cout, dzs)
            ldc = cats[L - 1].shape[-1]
            # dW^T[c][tap*3+co] = sum_pixels x[q][c] dzs[q][.]: a 1x1 "transposed" weight-gradient GEMM (S = x, B = dzs)
            gw = K.conv_geom(N, H, W, fc.cin, ldc, H, W, 32, 32, 1, 1, 0, True, K.IMPL_TC)
            nparts = K.conv_wgrad_parts(gw, Slice(cats[L - 1]), Slice(dzs), scratch)
            K.head_wgrad_scatter(scratch, nparts, fc.cin * 32, fc.cout, fc.cin, grads[fc.weight], True)
            K.bias_grad(dz4, 4, N * H * W, 3, grads[fc.bias])
            gd = K.conv_geom(N, H, W, 32, 32, H, W, fc.cin, ldc, 1, 1, 0, False, K.IMPL_TC)
            K.conv_forward(gd, Slice(dzs), None, self.head_wd, None, ACT_NONE, Slice(dcat), None, None)
        else:
            dzf = ws.get("dzf" + tag, (N, H, W, fc.dy_pad))      # channels 3.. stay zero (padding for the dgrad GEMM)
            K.tanh_bwd_combine(dout_nchw, dout_nhwc, sv["out"], dzf, fc.dy_pad, N, 3, H, W)
            fc.wgrad(Slice(cats[L - 1]), Slice(dz4), N, H, W, scratch, grads[fc.weight], accumulate)
            K.bias_grad(dz4, 4, N * H * W, 3, grads[fc.bias])
            fc.dgrad(Slice(dzf), N, H, W, Slice(dcat))
        dcats = {L - 1: dcat}

        image_grad = None
        side = self._side_stream(dcat.device)
        main = torch.cuda.current_stream() if side is not None else None
        scratch_side = ws.get("wgrad_scratch_side", (max(4 * max_w, 1 << 24),)) if side is not None else scratch
        # decoder blocks, last to first
        for j in range(L - 2, -1, -1):
            i, dprev, width = self._cat_layout(j)
            co = self.dec[j]
            oh, ow = hs[i - 1], wsz[i - 1]
            nl = self.dec_norm[j]
            si = st_idx[("dec", j)]
            dy = ws.get("dyd%d_%s" % (j, tag), (N, oh, ow, co))
            K.gn_bwd_reduce(Slice(dcats[j + 1], 0, co), Slice(cats[j + 1], 0, co), ACT_RELU, None, None, ACT_NONE,
                            drops[j] if j < 3 else None, sv["zd"][j], stats[si], N, oh * ow, co, dy, sums[si])
            K.gn_bwd_apply(dy, sv["zd"][j], stats[si], sums[si], nl.weight.detach(), N, oh * ow, co, grads[nl.weight],
                           grads[nl.bias])
            cv = self.dec_conv[j]
            # weight gradient and input gradient of a layer both only read dy: the weight gradients go to the side stream
            if side is not None:
                side.wait_stream(main)
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                cv.wgrad(Slice(cats[j]), Slice(dy), N, hs[i], wsz[i], scratch_side, grads[cv.weight], accumulate)
            dc = ws.get("dcat%d_%s" % (j, tag), tuple(cats[j].shape))
            cv.dgrad(Slice(dy), N, hs[i], wsz[i], Slice(dc), scratch=ws.get("splitk_main", (SPLITK_SCRATCH,)))
            dcats[j] = dc

        if side is not None:
            side.wait_stream(main)
        if on_stage is not None:
            # the decoder's weight gradients are on the side stream: issue the bucket's all-reduce behind them
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                on_stage("decoder", None)
        # Encoders, deepest level first, the branches level by level in lock step (the second branch on the side stream).
        # Levels >= DEEP_SPLIT hold most of an encoder's parameters (50 of 61 MB) and finish first: their gradient bucket is
        # handed to on_stage as soon as both branches have enqueued them, the shallow levels' bucket at the end.
        split = self.deep_split()
        dact_next = {name: None for name, _, _, _, _ in self.specs}
        dwarps = {}

        def ctx(e_idx):
            return torch.cuda.stream(side) if (side is not None and e_idx == 1) else contextlib.nullcontext()

        for e_idx, (name, _, _, _, warped_branch) in enumerate(self.specs):
            if warped_branch:
                with ctx(e_idx):
                    # gradient of all warped skip levels in ONE launch (zero fill of the scatter targets included)
                    lv = []
                    for i in range(min(4, L)):
                        j = L - 1 - i
                        _, dprev, _ = self._cat_layout(j)
                        c = self.enc[i]
                        off = dprev + e_idx * c
                        dwarps[i] = ws.get("dwarp%d_%s" % (i, tag), (N, hs[i], wsz[i], c))
                        lv.append(dict(dy=Slice(dcats[j], off, c), y=Slice(cats[j], off, c), mask=sv["mlv"][i], argk=sv["argk"][i],
                                       dx=dwarps[i], C=c, h=hs[i], w=wsz[i]))
                    K.warp_backward_levels(lv, sv["warps"], N, sv["K"], H0, W0, ACT_RELU, True)
        for i in range(L - 1, -1, -1):
            for e_idx, (name, _, _, _, warped_branch) in enumerate(self.specs):
                on_side = side is not None and e_idx == 1
                scratch_e = scratch_side if on_side else scratch
                with ctx(e_idx):
                    convs, norms = self.enc_conv[name], self.enc_norm[name]
                    j = L - 1 - i
                    _, dprev, _ = self._cat_layout(j)
                    c = self.enc[i]
                    HW = hs[i] * wsz[i]
                    off = dprev + e_idx * c
                    warped = warped_branch and i < 4
                    if warped:
                        skip_g, skip_a, skip_act = Slice(dwarps[i]), None, ACT_NONE
                    else:
                        skip_g, skip_a, skip_act = Slice(dcats[j], off, c), Slice(cats[j], off, c), ACT_RELU
                    norm = norms[i]
                    z = sv["z"][(name, i)]
                    dy = ws.get("dye_%s%d_%s" % (name, i, tag), (N, hs[i], wsz[i], c))
                    si = st_idx[(name, i)] if norm is not None else None
                    st = stats[si] if norm is not None else None
                    sm = sums[si] if norm is not None else None
                    if dact_next[name] is not None:
                        K.gn_bwd_reduce(Slice(dact_next[name]), Slice(sv["act"][(name, i)]), ACT_LEAKY, skip_g, skip_a, skip_act,
                                        None, z if norm is not None else None, st, N, HW, c, dy, sm)
                    else:
                        K.gn_bwd_reduce(skip_g, skip_a, skip_act, None, None, ACT_NONE, None, z if norm is not None else None,
                                        st, N, HW, c, dy, sm)
                    if norm is not None:
                        K.gn_bwd_apply(dy, z, st, sm, norm.weight.detach(), N, HW, c, grads[norm.weight], grads[norm.bias])
                    cv = convs[i]
                    if i == 0:
                        cv.wgrad(Slice(sv["xin"][name]), Slice(dy), N, H, W, scratch_e, grads[cv.weight], accumulate)
                        K.bias_grad(dy, c, N * HW, c, grads[cv.bias])
                        if need_image_grad and e_idx == 0:
                            # gradient w.r.t. the stem's input; channels 0..2 are the image (pose channels are data)
                            dxin = ws.get("dxin_%s_%s" % (name, tag), (N, H, W, cv.cin_pad))
                            cv.dgrad(Slice(dy), N, H, W, Slice(dxin), dx_channels=cv.cin_pad)
                            image_grad = torch.empty(N, 3, H, W, device=dxin.device)
                            K.nhwc_to_nchw(Slice(dxin, 0, 3), image_grad)
                    else:
                        cv.wgrad(Slice(sv["act"][(name, i - 1)]), Slice(dy), N, hs[i - 1], wsz[i - 1], scratch_e, grads[cv.weight], accumulate)
                        dact = ws.get("dact_%s%d_%s" % (name, i - 1, tag), (N, hs[i - 1], wsz[i - 1], self.enc[i - 1]))
                        cv.dgrad(Slice(dy), N, hs[i - 1], wsz[i - 1], Slice(dact),
                                 scratch=ws.get("splitk_side" if on_side else "splitk_main", (SPLITK_SCRATCH,)))
                        dact_next[name] = dact
                    if on_stage is not None and i == split and split > 0:
                        on_stage(name, "deep")
                    if on_stage is not None and i == 0:
                        on_stage(name, "shallow" if split > 0 else None)
        if side is not None:
            main.wait_stream(side)
        return image_grad


class _ForkedGeneratorEngine(GeneratorEngine):
    """GeneratorEngine.fork(): own `ws` / `saved`, everything else (layers, weight packs, streams) read from the parent."""

    def __getattr__(self, name):        # only reached when the attribute is not set on the fork itself
        if name == "_parent":
            raise AttributeError(name)
        return getattr(self._parent, name)


class DiscriminatorEngine:
    """Discriminator.forward / backward (models/networks.py:338-357): Conv(k4,s2,p0)+bias -> 3 normed
    Blocks -> Block(512,1,bn=False) -> Sigmoid -> Flatten."""

    def __init__(self, module):
        self.m = module
        self.ws = None
        self.layers_built = False
        self.saved = None

    def _build_layers(self):
        net = self.m.net
        self.convs = [ConvLayer(net[0].weight, net[0].bias, False, 4, 2, 0)]
        self.norms = [None]
        i = 1
        while hasattr(net[i], "net"):
            blk = net[i]
            self.convs.append(ConvLayer(blk.net[1].weight, None, False, 4, 2, 1))
            self.norms.append(NormLayer(blk.net[2].weight, blk.net[2].bias) if len(blk.net) > 2 else None)
            i += 1
        self.layers_built = True

    def _ensure(self, device):
        if not self.layers_built or self.convs[0].weight is not self.m.net[0].weight:
            self._build_layers()
        if self.ws is None or self.ws.device != device:
            self.ws = _Workspace(device)

    def input_buffer(self, M, H, W, device):
        """NHWC input buffer [M,H,W,ceil4(input_nc)] (zero padded channels) the caller fills by slices."""
        self._ensure(device)
        return self.ws.get("din_%d_%d_%d" % (M, H, W), (M, H, W, self.convs[0].cin_pad))

    def pack_weights(self):
        for c in self.convs:
            c.pack_forward()
        self.packed_version = _weights_version(self.m)

    def _need_repack(self, repack):
        if repack is None:
            v = _weights_version(self.m)
            return v is None or v != getattr(self, "packed_version", -1) or self.convs[0].w_fwd is None
        return bool(repack)

    def forward(self, din, repack=True, probs=False):
        """din = input_buffer() filled by the caller.  Returns logits [M, OH*OW] (pre-sigmoid), or the sigmoid
        probabilities if probs=True (what Discriminator.forward returns in the reference)."""
        M, H, W, _ = din.shape
        self._ensure(din.device)
        ws = self.ws
        if self._need_repack(repack):
            self.pack_weights()
        tag = "%d_%d_%d" % (M, H, W)
        nl = len(self.convs)
        stats = ws.get("dstats" + tag, (nl, M, 2), torch.float64, zero=True)
        sv = {"M": M, "H": H, "W": W, "tag": tag, "din": din, "stats": stats, "z": [], "act": [], "hw": []}
        x, h, w = din, H, W
        for i, cv in enumerate(self.convs):
            oh, ow = cv.out_hw(h, w)
            last = i == nl - 1
            if i == 0:
                act = ws.get("dact0_" + tag, (M, oh, ow, cv.cout))
                cv.forward(Slice(x), M, h, w, Slice(act), ACT_LEAKY)
                sv["z"].append(None)
                sv["act"].append(act)
            elif not last:
                z = ws.get("dz%d_%s" % (i, tag), (M, oh, ow, cv.cout))
                cv.forward(Slice(x), M, h, w, Slice(z), ACT_NONE, stats[i], scratch=ws.get("splitk_main", (SPLITK_SCRATCH,)))
                act = ws.get("dact%d_%s" % (i, tag), (M, oh, ow, cv.cout))
                nm = self.norms[i]
                K.gn_apply(z, stats[i], nm.weight.detach(), nm.bias.detach(), None, M, oh * ow, cv.cout, act, ACT_LEAKY)
                sv["z"].append(z)
                sv["act"].append(act)
            else:
                logits = torch.empty(M, oh * ow * cv.cout, device=din.device)
                cv.forward(Slice(x), M, h, w, Slice(logits.view(M, oh, ow, cv.cout)), ACT_SIGMOID if probs else ACT_NONE)
                sv["logits_hw"] = (oh, ow)
                act = logits
            sv["hw"].append((h, w))
            x, h, w = act, oh, ow
        self.saved = sv
        return x

    def dlogits_buffer(self, M, J):
        """Zero-padded gradient buffer [M*J, dy_pad] for the logits (channel 0 carries the gradient)."""
        return self.ws.get("dlog_%d_%d" % (M, J), (M * J, self.convs[-1].dy_pad))

    def backward(self, dlogits4, grads=None, need_input_grad=False):
        """dlogits4: dlogits_buffer() filled by the caller (gradient w.r.t. the logits in channel 0, zeros elsewhere).
        grads: dict parameter -> gradient tensor (None => no weight gradients, G-step use).
        Returns the gradient w.r.t. the NHWC input buffer if need_input_grad."""
        sv = self.saved
        ws = self.ws
        M, tag = sv["M"], sv["tag"]
        nl = len(self.convs)
        stats = sv["stats"]
        sums = ws.get("dsums" + tag, tuple(stats.shape), torch.float64, zero=True)
        max_w = max(c.taps * c.cin_pad * c.cout_pad for c in self.convs)
        scratch = ws.get("wgrad_scratch", (max(4 * max_w, 1 << 24),)) if grads is not None else None
        dummy = ws.get("dummy_gb", (2,))
        dy = dlogits4
        din_grad = None
        side = None
        if grads is not None and dlogits4.is_cuda and STREAMS and os.environ.get("PTK_STREAMS", "1") != "0":
            side = getattr(self, "_side", None)
            if side is None or side.device != dlogits4.device:
                side = self._side = torch.cuda.Stream(device=dlogits4.device)
        main = torch.cuda.current_stream() if side is not None else None
        for i in range(nl - 1, -1, -1):
            cv = self.convs[i]
            h, w = sv["hw"][i]
            oh, ow = cv.out_hw(h, w)
            x = sv["din"] if i == 0 else sv["act"][i - 1]
            dy_s = Slice(dy.view(M, oh, ow, -1))
            if grads is not None:
                # weight / bias gradients only read dy: side stream, concurrent with the input-gradient chain
                if side is not None:
                    side.wait_stream(main)
                with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                    cv.wgrad(Slice(x), dy_s, M, h, w, scratch, grads[cv.weight])
                    if cv.bias is not None:
                        K.bias_grad(dy, dy_s.ld, M * oh * ow, cv.cout, grads[cv.bias])
            if i == 0:
                if need_input_grad:
                    din_grad = ws.get("din_grad" + tag, tuple(sv["din"].shape))
                    cv.dgrad(dy_s, M, h, w, Slice(din_grad), dx_channels=cv.cin_pad)
                break
            # gradient w.r.t. the previous block's activated output, then through its activation / norm
            pc = self.convs[i - 1]
            dact = ws.get("ddact%d_%s" % (i - 1, tag), (M, h, w, pc.cout))
            cv.dgrad(dy_s, M, h, w, Slice(dact), scratch=ws.get("splitk_main", (SPLITK_SCRATCH,)))
            dprev = ws.get("ddy%d_%s" % (i - 1, tag), (M, h, w, pc.cout))
            nm = self.norms[i - 1]
            if nm is not None:
                K.gn_bwd_reduce(Slice(dact), Slice(sv["act"][i - 1]), ACT_LEAKY, None, None, ACT_NONE, None, sv["z"][i - 1],
                                stats[i - 1], M, h * w, pc.cout, dprev, sums[i - 1])
                K.gn_bwd_apply(dprev, sv["z"][i - 1], stats[i - 1], sums[i - 1], nm.weight.detach(), M, h * w, pc.cout,
                               grads[nm.weight] if grads is not None else dummy[0:1],
                               grads[nm.bias] if grads is not None else dummy[1:2])
            else:
                K.gn_bwd_reduce(Slice(dact), Slice(sv["act"][i - 1]), ACT_LEAKY, None, None, ACT_NONE, None, None, None, M,
                                h * w, pc.cout, dprev, None)
            dy = dprev
        if side is not None:
            main.wait_stream(side)
        return din_grad
