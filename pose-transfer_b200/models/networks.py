"""Drop-in for the reference's ``models/networks.py`` (src_deformable/models/networks.py:130-357).

Same class names, constructor signatures, sub-module tree and therefore the same ``state_dict`` keys and
shapes (checkpoint ABI), but ``forward`` runs the hand-written sm_100a kernels through
``engine.GeneratorEngine`` / ``engine.DiscriminatorEngine``.  The leaf ``nn.Conv2d`` / ``nn.InstanceNorm3d``
objects only OWN parameters; their own forward is never called.  There is no eager fallback: CPU tensors
raise.
"""
import torch
import torch.nn as nn

from ..engine import DiscriminatorEngine, GeneratorEngine
from ..kernels import Slice
from ..utils import pose_utils
from ..utils.pose_transform import AffineTransformLayer  # noqa: F401  (re-exported like the reference)


def print_network(net):
    """models/networks.py:18-24."""
    num_params = sum(p.numel() for p in net.parameters())
    print(net)
    print('Total number of parameters: %d' % num_params)


def xavier_weights_init(m):
    """models/networks.py:26-31 (not applied by DeformablePose_GAN, pose_gan.py:62-65)."""
    classname = m.__class__.__name__
    if classname.find('Conv') == 0 and hasattr(m, 'weight'):
        nn.init.xavier_uniform_(m.weight.data, gain=1)


def gaussian_weights_init(m):
    classname = m.__class__.__name__
    if classname.find('Conv') == 0 and hasattr(m, 'weight'):
        m.weight.data.normal_(0.0, 0.02)


class Flatten(nn.Module):
    def forward(self, input):
        return input.view(input.size(0), -1)


class Cropping2D(nn.Module):
    def __init__(self, crop_size):
        super(Cropping2D, self).__init__()
        self.crop_size = crop_size

    def forward(self, input):
        return input[:, :, self.crop_size:-self.crop_size, self.crop_size:-self.crop_size]


def _bump_weights_version(module, incompatible_keys):
    """load_state_dict post hook: the engines' GEMM-layout weight copies are stale."""
    module._ptk_weights_version = getattr(module, "_ptk_weights_version", 0) + 1


def _require_cuda(t, who):
    """The product has no CPU path: fail loudly (the CPU test-suite swaps this guard out together with the kernels)."""
    if not t.is_cuda:
        raise RuntimeError("%s: CUDA tensors required (no CPU fallback)" % who)


def _no_eager(name):
    raise RuntimeError("%s.forward: this module only owns parameters; the computation runs inside the fused "
                       "CUDA schedule of its parent network (no eager fallback)" % name)


class Block(nn.Module):
    """Parameter container mirroring models/networks.py:142-172 (module indices = state_dict keys)."""

    def __init__(self, input_nc, output_nc, down=True, bn=True, dropout=False, leaky=True):
        super(Block, self).__init__()
        self.net = self.build_net(input_nc, output_nc, down, bn, dropout, leaky)

    def build_net(self, input_nc, output_nc, down=True, bn=True, dropout=False, leaky=True):
        model = [nn.LeakyReLU(0.2) if leaky else nn.ReLU()]
        if down:
            model.append(nn.Conv2d(input_nc, output_nc, kernel_size=4, stride=2, padding=1, bias=False))
        else:
            model.append(nn.ConvTranspose2d(input_nc, output_nc, kernel_size=4, stride=2, bias=False))
            model.append(Cropping2D(1))
        if bn:
            model.append(nn.InstanceNorm3d(1, eps=1e-3, affine=True, track_running_stats=False))
        if dropout:
            model.append(nn.Dropout2d())
        return nn.ModuleList(model)

    def forward(self, input):
        _no_eager("Block")


class encoder(nn.Module):
    """models/networks.py:175-202."""

    def __init__(self, input_nc, nfilters_enc):
        super(encoder, self).__init__()
        self.input_nc = input_nc
        self.nfilters_enc = nfilters_enc
        self.net = self.build_net(input_nc, nfilters_enc)

    def build_net(self, input_nc, nfilters_enc):
        model = []
        for i, nf in enumerate(nfilters_enc):
            if i == 0:
                model.append(nn.Conv2d(input_nc, nf, kernel_size=3, padding=1, bias=True))
            elif i == len(nfilters_enc) - 1:
                model.append(Block(nfilters_enc[i - 1], nf, bn=False))
            else:
                model.append(Block(nfilters_enc[i - 1], nf))
        return nn.ModuleList(model)

    def forward(self, input):
        _no_eager("encoder")


class decoder(nn.Module):
    """models/networks.py:204-250."""

    def __init__(self, nfilters_dec, nfilters_enc, num_skips=1):
        super(decoder, self).__init__()
        self.num_skips = num_skips
        self.nfilters_dec = nfilters_dec
        self.nfilters_enc = nfilters_enc
        self.net = self.build_net(nfilters_dec)

    def build_net(self, nfilters_dec):
        model_dec = []
        for i, nf in enumerate(nfilters_dec):
            if i == 0:
                model_dec.append(Block(self.num_skips * self.nfilters_enc[-1], nf, down=False, leaky=False, dropout=True))
            elif i == len(nfilters_dec) - 1:
                model_dec.append(nn.ReLU())
                model_dec.append(nn.Conv2d(self.num_skips * self.nfilters_enc[-(i + 1)] + nfilters_dec[i - 1], nf,
                                           kernel_size=3, padding=1, bias=True))
            elif 0 < i < 3:
                model_dec.append(Block(self.num_skips * self.nfilters_enc[-(i + 1)] + nfilters_dec[i - 1], nf, down=False,
                                       leaky=False, dropout=True))
            else:
                model_dec.append(Block(self.num_skips * self.nfilters_enc[-(i + 1)] + nfilters_dec[i - 1], nf, down=False,
                                       leaky=False))
        model_dec.append(nn.Tanh())
        return nn.ModuleList(model_dec)

    def forward(self, skips):
        _no_eager("decoder")


class _GeneratorFn(torch.autograd.Function):
    """Makes the fused schedule differentiable for external callers (main.py / test.py style use).  `engine` is the
    execution context that holds this call's activations (the module's own engine, or one fork per stack of the stacked
    generator).  Gradients: every parameter, and the IMAGE channels 0..2 of `inp` (what a previous stack's output feeds,
    models/networks.py:320-323); the pose heat-map channels are data and get zeros."""

    @staticmethod
    def forward(ctx, gen, engine, inp, warps, masks, *params):
        ctx.engine = engine
        ctx.params = params
        out = engine.forward(inp, warps, masks, drop=gen._next_drop())
        ctx.token = engine.saved
        ctx.inp_shape = tuple(inp.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        engine = ctx.engine
        if engine.saved is not ctx.token:
            raise RuntimeError("the generator was run again before this backward: its saved activations are gone")
        # fresh contiguous tensors (zeros_like would inherit the GEMM-layout strides of arena-backed parameters)
        grads = {p: torch.zeros(p.shape, device=p.device) for p in ctx.params}
        need_inp = ctx.needs_input_grad[2]
        ig = engine.backward(grads, dout_nchw=dout.contiguous(), need_image_grad=need_inp)
        dinp = None
        if need_inp:
            dinp = torch.zeros(ctx.inp_shape, device=dout.device)
            dinp[:, :3] = ig
        return (None, None, dinp, None, None) + tuple(grads[p] if p.requires_grad else None for p in ctx.params)


class Deformable_Generator(nn.Module):
    """models/networks.py:252-288."""

    def __init__(self, input_nc, pose_dim, image_size, nfilters_enc, nfilters_dec, warp_skip, use_input_pose=True):
        super(Deformable_Generator, self).__init__()
        self.input_nc = input_nc
        # 'none' (opts.py:60) never equals 'None' (networks.py:257): num_skips is always 2 in the reference
        self.num_skips = 1 if warp_skip == 'None' else 2
        self.warp_skip = warp_skip
        self.pose_dim = pose_dim
        self.nfilters_dec = nfilters_dec
        self.nfilters_enc = nfilters_enc
        self.image_size = image_size
        self.use_input_pose = use_input_pose
        if warp_skip != 'mask' or not use_input_pose:
            raise NotImplementedError("only warp_skip='mask' with use_input_pose=True is on the B200 hot path")
        self.encoder_app = encoder(input_nc - self.pose_dim, nfilters_enc)
        self.encoder_pose = encoder(self.pose_dim, nfilters_enc)
        self.decoder = decoder(nfilters_dec, nfilters_enc, self.num_skips)
        self.engine = GeneratorEngine(self)
        self._drop_queue = None
        self._ptk_weights_version = 0
        self.register_load_state_dict_post_hook(_bump_weights_version)

    def set_dropout_noise(self, drops):
        """Test hook: the next forward uses these three [N,512,1,1] noise tensors instead of drawing."""
        self._drop_queue = drops

    def _next_drop(self):
        d, self._drop_queue = self._drop_queue, None
        return d

    def forward(self, input, warps, masks):
        _require_cuda(input, "Deformable_Generator")
        params = tuple(self.parameters())
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in params)):
            return _GeneratorFn.apply(self, self.engine, input, warps, masks, *params)
        return self.engine.forward(input, warps, masks, drop=self._next_drop())


class Generator(nn.Module):
    """Drop-in for src_baseline/models/networks.py:238-254: the same U-Net with ONE encoder over the whole
    input and no warp (num_skips = 1), on the same kernels / engine.  state_dict keys: encoder.net.*, decoder.net.*."""

    def __init__(self, input_nc, nfilters_enc, nfilters_dec, num_skips=1, warp_skip=False, use_input_pose=True):
        super(Generator, self).__init__()
        if num_skips != 1:
            raise NotImplementedError("src_baseline Generator: only num_skips=1 (the reference's only use) is supported")
        self.input_nc = input_nc
        self.num_skips = num_skips
        self.nfilters_dec = nfilters_dec
        self.nfilters_enc = nfilters_enc
        self.image_size = None
        self.encoder = encoder(input_nc, nfilters_enc)
        self.decoder = decoder(nfilters_dec, nfilters_enc, num_skips)
        self.engine = GeneratorEngine(self)
        self._drop_queue = None
        self._ptk_weights_version = 0
        self.register_load_state_dict_post_hook(_bump_weights_version)

    def set_dropout_noise(self, drops):
        self._drop_queue = drops

    def _next_drop(self):
        d, self._drop_queue = self._drop_queue, None
        return d

    def forward(self, input):
        _require_cuda(input, "Generator")
        params = tuple(self.parameters())
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in params)):
            return _GeneratorFn.apply(self, self.engine, input, None, None, *params)
        return self.engine.forward(input, None, None, drop=self._next_drop())


class Stacked_Generator(nn.Module):
    """models/networks.py:290-327: the same Deformable_Generator applied `num_stacks` times, stack i reading
    [previous output | pose_{i-1} | pose_i] with its own warps / masks.  One engine context per stack (shared layers and
    weight packs, private activations) keeps every stack's forward alive for the backward pass, which walks the stacks in
    reverse and hands the gradient of the image channels from stack i to stack i-1."""

    def __init__(self, input_nc, num_stacks, image_size, pose_dim, nfilters_enc, nfilters_dec, warp_skip=False,
                 use_input_pose=True):
        super(Stacked_Generator, self).__init__()
        self.input_nc = input_nc
        self.num_stacks = num_stacks
        self.nfilters_dec = nfilters_dec
        self.nfilters_enc = nfilters_enc
        self.use_input_pose = use_input_pose
        self.pose_dim = pose_dim
        self.image_size = image_size
        self.generator = self._make_generator(input_nc, pose_dim, image_size, nfilters_enc, nfilters_dec, warp_skip, use_input_pose)
        self._contexts = []

    def _make_generator(self, input_nc, pose_dim, image_size, nfilters_enc, nfilters_dec, warp_skip, use_input_pose):
        return Deformable_Generator(input_nc, pose_dim, image_size, nfilters_enc, nfilters_dec, warp_skip, use_input_pose)

    @property
    def engine(self):
        return self.generator.engine

    def contexts(self, n):
        """Engine contexts of the first n stacks (stack 0 = the generator's own engine)."""
        while len(self._contexts) < n:
            self._contexts.append(self.generator.engine if not self._contexts else self.generator.engine.fork())
        return self._contexts[:n]

    def stack_input(self, i, input, target_pose, prev_out):
        """Pieces of stack i's input (models/networks.py:312-323), never concatenated in memory."""
        P = self.pose_dim
        if i == 0:
            return [(input, 0, 3 + P), (target_pose, 0, P)]
        return [(prev_out, 0, 3), (target_pose, (i - 1) * P, P), (target_pose, i * P, P)]

    def run_stacks(self, input, target_pose, target_warps, target_masks, drops=None, d_input=None, repack=None):
        """Trainer path (no autograd graph): returns the list of stack outputs; d_input receives the LAST output."""
        S = self.num_stacks
        ctxs = self.contexts(S)
        outs = []
        for i, eng in enumerate(ctxs):
            pieces = self.stack_input(i, input, target_pose, outs[-1] if outs else None)
            w = target_warps[:, i].contiguous() if target_warps is not None else None
            m = target_masks[:, i].contiguous() if target_masks is not None else None
            outs.append(eng.forward(pieces, w, m, drop=drops[i] if drops is not None else None,
                                    repack=repack if i == 0 else False, d_input=d_input if i == S - 1 else None))
        return outs

    def backward_stacks(self, grads, dout_nchw, dout_nhwc, on_stage=None):
        """Gradients of a loss on the last stack's output w.r.t. the shared parameters (summed over the stacks)."""
        S = self.num_stacks
        ctxs = self.contexts(S)
        g, g2 = dout_nchw, dout_nhwc
        for i in range(S - 1, -1, -1):
            g = ctxs[i].backward(grads, dout_nchw=g, dout_nhwc=g2, on_stage=on_stage if i == 0 else None,
                                 accumulate=i < S - 1, need_image_grad=i > 0)
            g2 = None

    def forward(self, input, target_pose, target_warps=None, target_masks=None):
        _require_cuda(input, "Stacked_Generator")
        gen = self.generator
        params = tuple(gen.parameters())
        track = torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in params))
        outputs = []
        for i, eng in enumerate(self.contexts(self.num_stacks)):
            w = target_warps[:, i].contiguous() if target_warps is not None else None
            m = target_masks[:, i].contiguous() if target_masks is not None else None
            if track:
                # autograd needs ONE input tensor per stack: materialise the concatenation (module-surface use only)
                inp = torch.cat([t[:, c0:c0 + C] for t, c0, C in self.stack_input(i, input, target_pose, outputs[-1] if outputs else None)], 1)
                out = _GeneratorFn.apply(gen, eng, inp, w, m, *params)
            else:
                out = eng.forward(self.stack_input(i, input, target_pose, outputs[-1] if outputs else None), w, m,
                                  drop=gen._next_drop(), repack=True if i == 0 else False)
            outputs.append(out)
        return outputs


class Stacked_Baseline_Generator(Stacked_Generator):
    """Drop-in for src_baseline/models/networks.py:255-298: the same stacking over the single-encoder, un-warped Generator;
    called as gen(input, interpol_pose)."""

    def __init__(self, input_nc, num_stacks, pose_dim, nfilters_enc, nfilters_dec, num_skips=1, warp_skip=False,
                 use_input_pose=True):
        self._num_skips = num_skips
        super(Stacked_Baseline_Generator, self).__init__(input_nc, num_stacks, None, pose_dim, nfilters_enc, nfilters_dec,
                                                         warp_skip, use_input_pose)

    def _make_generator(self, input_nc, pose_dim, image_size, nfilters_enc, nfilters_dec, warp_skip, use_input_pose):
        return Generator(input_nc, nfilters_enc, nfilters_dec, self._num_skips, warp_skip, use_input_pose)


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disc, x, *params):
        ctx.disc, ctx.params = disc, params
        probs = disc._run(x)
        ctx.save_for_backward(probs)
        ctx.x_shape = x.shape
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        from .. import kernels as K
        disc = ctx.disc
        probs, = ctx.saved_tensors
        M, J = probs.shape
        # sigmoid' (tiny [M,J] tensor; part of the autograd glue, not of the step's hot path)
        dlog = (dprobs * probs * (1 - probs)).reshape(M * J, 1)
        dlog4 = disc.engine.dlogits_buffer(M, J)
        dlog4[:, :1] = dlog
        grads = {p: torch.zeros(p.shape, device=p.device) for p in ctx.params}
        need_x = ctx.needs_input_grad[1]
        din_grad = disc.engine.backward(dlog4, grads, need_input_grad=need_x)
        dx = None
        if need_x:
            N, C, H, W = ctx.x_shape
            dx = torch.empty(N, C, H, W, device=probs.device)
            K.nhwc_to_nchw(Slice(din_grad, 0, C), dx)
        return (None, dx) + tuple(grads[p] if p.requires_grad else None for p in ctx.params)


class Discriminator(nn.Module):
    """models/networks.py:329-357."""

    def __init__(self, input_nc, warp_skip=False, use_input_pose=True, checkMode=0, baseline_tree=False):
        super(Discriminator, self).__init__()
        self.input_nc = input_nc
        self.use_input_pose = use_input_pose
        self.warp_skip = warp_skip
        self.checkMode = checkMode
        # the two trees reduce the PatchGAN differently under --checkMode: src_deformable keeps Block(128, 256) and ends with
        # Block(256, 1) (networks.py:341-351), src_baseline ends with Block(128, 1) (src_baseline/models/networks.py:312-319)
        self.baseline_tree = baseline_tree
        self.net = self.build_net()
        self.engine = DiscriminatorEngine(self)
        self._ptk_weights_version = 0
        self.register_load_state_dict_post_hook(_bump_weights_version)

    def build_net(self):
        model = [nn.Conv2d(self.input_nc, 64, kernel_size=4, stride=2), Block(64, 128)]
        if self.checkMode == 0:
            model += [Block(128, 256), Block(256, 512), Block(512, 1, bn=False)]
        elif self.baseline_tree:
            model.append(Block(128, 1, bn=False))
        else:
            model += [Block(128, 256), Block(256, 1, bn=False)]
        model.append(nn.Sigmoid())
        model.append(Flatten())
        return nn.Sequential(*model)

    def _run(self, x):
        from .. import kernels as K
        M, C, H, W = x.shape
        din = self.engine.input_buffer(M, H, W, x.device)
        K.gather_nhwc([(x.contiguous(), 0, C, 0)], Slice(din, 0, din.shape[-1]), din.shape[-1])
        return self.engine.forward(din, probs=True)

    def forward(self, input):
        _require_cuda(input, "Discriminator")
        params = tuple(self.parameters())
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in params)):
            return _DiscriminatorFn.apply(self, input, *params)
        return self._run(input)
