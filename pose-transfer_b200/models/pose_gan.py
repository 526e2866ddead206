"""Drop-in for the reference's ``models/pose_gan.py`` (src_deformable/models/pose_gan.py:11-224):
``DeformablePose_GAN`` with the same constructor (an ``opts`` Namespace), attributes (``gen``, ``disc``,
``gen_opt``, ``disc_opt``, ``content_model``) and methods (``gen_update``, ``dis_update``, ``nn_loss``,
``resume``, ``save``), executing one D step / one G step as a hand-scheduled sequence of sm_100a kernels.

Differences that do not change any returned value or parameter update:
  * ``dis_update`` does not back-propagate into the generator (the reference's G gradients from that
    backward are discarded by ``gen.zero_grad()`` before they are ever used);
  * ``gen_update`` computes no discriminator / VGG weight gradients (never used by the reference);
  * losses are reduced by one fused kernel each and read back with a single device->host copy.
Data parallelism (absent from the reference): if ``torch.distributed`` is initialised, gradients are
summed over ranks with NCCL on the flat gradient arena and the 1/world scale is folded into Adam;
``opt.batch_size`` is the PER-RANK batch.
"""
import contextlib
import os

import torch
import torch.nn as nn
from torchvision.models import vgg19

from .. import kernels as K
from ..kernels import Slice
from ..utils import pose_utils
from .networks import (Deformable_Generator, Discriminator, Generator, Stacked_Baseline_Generator, Stacked_Generator,  # noqa: F401
                       xavier_weights_init)


class ParamArena:
    """All parameters of a network as views into ONE flat fp32 buffer (plus flat grad / Adam state), so the
    optimiser step is one kernel launch and the gradient all-reduce is one NCCL call."""

    def __init__(self, module):
        self.module = module
        self.params = [p for p in module.parameters()]
        self.bind()

    def bump(self):
        """Weights changed: the engines' GEMM-layout copies are stale (every sub-module that owns an engine is told:
        a Stacked_Generator wraps the Deformable_Generator whose engine holds the packs)."""
        for m in self.module.modules():
            if hasattr(m, "_ptk_weights_version"):
                m._ptk_weights_version += 1

    @staticmethod
    def _gemm_master(p):
        """Conv / ConvTranspose weights [A][B][kh][kw] with A % 32 == 0 live in the arena in GEMM layout
        [tap][A][B_pad] (B_pad = B rounded up to 32): that is at once the K-major tensor-core operand of one
        direction and the layout the weight gradient is produced in, so the optimiser step needs no unpack and
        only ONE derived layout (a per-tap transpose).  The nn.Parameter is a strided VIEW of that storage with the
        checkpoint shape, so state_dict()/load_state_dict() keep the reference's ABI."""
        return p.dim() == 4 and p.shape[0] % 32 == 0 and p.shape[2] * p.shape[3] <= 16

    def bind(self):
        self.bump()
        dev = self.params[0].device
        total = 0
        self.offsets, self.lengths = [], []
        for p in self.params:
            total = (total + 3) // 4 * 4          # keep every tensor 16-byte aligned
            self.offsets.append(total)
            if self._gemm_master(p):
                A, B, kh, kw = p.shape
                n = kh * kw * A * ((B + 31) // 32 * 32)
            else:
                n = p.numel()
            self.lengths.append(n)
            total += n
        self.total = (total + 3) // 4 * 4
        old_m, old_v = getattr(self, "exp_avg", None), getattr(self, "exp_avg_sq", None)
        self.flat = torch.zeros(self.total, device=dev)
        self.grad = torch.zeros(self.total, device=dev)
        self.exp_avg = torch.zeros(self.total, device=dev)
        self.exp_avg_sq = torch.zeros(self.total, device=dev)
        if old_m is not None and old_m.numel() == self.total:
            # re-bind after e.g. module.to(...): same parameters in the same order => same arena layout; the optimiser state
            # moves with the weights (FlatAdam keeps its step count, so the bias correction stays consistent)
            self.exp_avg.copy_(old_m)
            self.exp_avg_sq.copy_(old_v)
        self.grads = {}
        with torch.no_grad():
            for p, off, n in zip(self.params, self.offsets, self.lengths):
                if self._gemm_master(p):
                    A, B, kh, kw = p.shape
                    Bp = n // (kh * kw * A)
                    strides = (Bp, 1, kw * A * Bp, A * Bp)
                    view = self.flat[off:off + n].as_strided(tuple(p.shape), strides)
                    g = self.grad[off:off + n].as_strided(tuple(p.shape), strides)
                    p._ptk_master = (self.flat[off:off + n], self.grad[off:off + n], A, Bp, kh * kw)
                else:
                    view = self.flat[off:off + n].view(p.shape)
                    g = self.grad[off:off + n].view(p.shape)
                    p._ptk_master = None
                view.copy_(p.data)
                p.data = view
                p.grad = g
                self.grads[p] = g

    def check(self):
        """Re-flatten if some external code replaced parameter storage (e.g. module.to(...))."""
        p0, pl = self.params[0], self.params[-1]
        ok = (p0.data_ptr() == self.flat.data_ptr() + 4 * self.offsets[0] and
              pl.data_ptr() == self.flat.data_ptr() + 4 * self.offsets[-1])
        if not ok:
            self.bind()
        for p in self.params:     # someone may have set .grad = None (zero_grad(set_to_none=True))
            if p.grad is None or p.grad.data_ptr() != self.grads[p].data_ptr():
                p.grad = self.grads[p]

    def zero_grad(self):
        K.fill(self.grad, 0.0)

    def segment_range(self, submodule):
        """[lo, hi) of the flat arenas covering every parameter of `submodule` (parameters() enumerates sub-module by
        sub-module, so the range is contiguous; 16-byte aligned at both ends)."""
        ids = {id(p) for p in submodule.parameters()}
        idx = [i for i, p in enumerate(self.params) if id(p) in ids]
        assert idx and idx == list(range(idx[0], idx[-1] + 1)), "sub-module parameters are not contiguous in the arena"
        lo = self.offsets[idx[0]]
        hi = self.offsets[idx[-1] + 1] if idx[-1] + 1 < len(self.offsets) else self.total
        return lo, hi

    def segment_range_of(self, modules):
        """[lo, hi) covering every parameter of a list of sub-modules that are adjacent in the arena."""
        rs = [self.segment_range(m) for m in modules if any(True for _ in m.parameters())]
        rs.sort()
        assert rs and all(a[1] == b[0] for a, b in zip(rs, rs[1:])), "sub-modules are not adjacent in the arena"
        return rs[0][0], rs[-1][1]

    def segment(self, submodule):
        lo, hi = self.segment_range(submodule)
        return self.grad[lo:hi]


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas=(0.5, 0.999)) semantics (pose_gan.py:49-51) on a ParamArena."""

    def __init__(self, arena, lr, betas=(0.5, 0.999), eps=1e-8):
        super().__init__(arena.params, dict(lr=lr, betas=betas, eps=eps))
        self.arena = arena
        self.steps = 0
        self.grad_scale = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        self.begin_step()
        self.step_range(0, self.arena.total)
        self.end_step()

    # The same update issued range by range (a sub-network's bucket as soon as its gradients are final):
    def begin_step(self):
        self.steps += 1

    @torch.no_grad()
    def step_range(self, lo, hi):
        g = self.param_groups[0]
        a = self.arena
        K.adam_step(a.flat[lo:hi], a.grad[lo:hi], a.exp_avg[lo:hi], a.exp_avg_sq[lo:hi], g["lr"], g["betas"][0], g["betas"][1],
                    g["eps"], self.steps, self.grad_scale)

    def end_step(self):
        self.arena.bump()

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()


class _NNLossFn(torch.autograd.Function):
    """nn_loss on feature tensors through ptk_nnloss_features_* (gradient w.r.t. the prediction only is ever used)."""

    @staticmethod
    def forward(ctx, pred, gt, area):
        if not pred.is_cuda:
            raise RuntimeError("nn_loss: CUDA tensors required (no CPU fallback)")
        pred, gt = pred.detach().float().contiguous(), gt.detach().float().contiguous()
        N, C, H, W = pred.shape
        loss = torch.zeros(1, device=pred.device)
        argmin = torch.empty(N, H, W, dtype=torch.uint8, device=pred.device)
        K.nnloss_features_forward(pred, gt, area, 1.0, loss, argmin)
        ctx.save_for_backward(pred, gt, argmin)
        ctx.area = area
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        pred, gt, argmin = ctx.saved_tensors
        dpred = torch.empty_like(pred)
        K.nnloss_features_backward(pred, gt, argmin, ctx.area, 1.0, dpred)
        return dpred * gout, None, None


class DeformablePose_GAN(nn.Module):
    def __init__(self, opt):
        super(DeformablePose_GAN, self).__init__()
        # adding extra layers for larger image size (pose_gan.py:17-18)
        nfilters_decoder = (512, 512, 512, 256, 128, 3) if max(opt.image_size) < 256 else (512, 512, 512, 512, 256, 128, 3)
        nfilters_encoder = (64, 128, 256, 512, 512, 512) if max(opt.image_size) < 256 else (64, 128, 256, 512, 512, 512, 512)
        input_nc = 3 + 2 * opt.pose_dim if opt.use_input_pose else 3 + opt.pose_dim
        self.batch_size = opt.batch_size
        self.num_stacks = opt.num_stacks
        self.pose_dim = opt.pose_dim
        self.image_size = tuple(opt.image_size)
        if opt.gen_type == 'stacked':
            # the stacked generator is a composition of the same Deformable_Generator (one engine context per
            # stack); forward (test.py) and both training branches (pose_gan.py:72-77,120-125) run on the B200 path.
            self.gen = Stacked_Generator(input_nc, opt.num_stacks, opt.image_size, opt.pose_dim, nfilters_encoder,
                                         nfilters_decoder, opt.warp_skip, use_input_pose=opt.use_input_pose)
            pretrained_gen_path = '../exp/' + 'full_' + opt.dataset + '/models/gen_090.pkl'   # pose_gan.py:31-32
            try:
                self.gen.generator.load_state_dict(torch.load(pretrained_gen_path))
                print("Loaded generator from pretrained model ")
            except (FileNotFoundError, OSError):
                print("No pretrained generator at %s -- keeping default initialisation" % pretrained_gen_path)
        elif opt.gen_type == 'baseline':
            self.gen = Deformable_Generator(input_nc, self.pose_dim, opt.image_size, nfilters_encoder, nfilters_decoder,
                                            opt.warp_skip, use_input_pose=opt.use_input_pose)
        else:
            raise Exception('Invalid gen_type')
        self.disc = Discriminator(input_nc + 3, use_input_pose=opt.use_input_pose)
        # the reference unconditionally loads this checkpoint (pose_gan.py:40-42)
        pretrained_disc_path = '../exp/' + 'full_' + opt.dataset + '/models/disc_090.pkl'
        try:
            self.disc.load_state_dict(torch.load(pretrained_disc_path))
            print("Loaded discriminator from pretrained model ")
        except (FileNotFoundError, OSError):
            print("No pretrained discriminator at %s -- keeping default initialisation" % pretrained_disc_path)

        self._finish_init(opt)

    def _finish_init(self, opt):
        """Everything after the networks exist: rank discovery, content model, parameter arenas, optimisers."""
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size()
            self.rank = torch.distributed.get_rank()
            if torch.cuda.is_available():
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", self.rank % max(torch.cuda.device_count(), 1))))
        else:
            self.world, self.rank = 1, 0

        self.content_loss_layer = opt.content_loss_layer
        self.nn_loss_area_size = opt.nn_loss_area_size
        if self.content_loss_layer != 'none':
            try:
                self.content_model = vgg19(pretrained=True)
            except Exception as e:  # offline: no ImageNet weights reachable
                print("vgg19(pretrained=True) unavailable (%s) -- using torchvision default init" % type(e).__name__)
                self.content_model = vgg19(weights=None)
        self.gen.cuda()
        self.disc.cuda()
        self._nn_loss_area_size = opt.nn_loss_area_size
        self.ll_loss_criterion = torch.nn.L1Loss()

        lr = opt.learning_rate
        self.gen_arena = ParamArena(self.gen)
        self.disc_arena = ParamArena(self.disc)
        self.disc_opt = FlatAdam(self.disc_arena, lr=lr, betas=(0.5, 0.999))
        self.gen_opt = FlatAdam(self.gen_arena, lr=lr, betas=(0.5, 0.999))
        self.gen_opt.grad_scale = self.disc_opt.grad_scale = 1.0 / self.world
        if self.world > 1:   # identical replicas: rank 0's weights win
            torch.distributed.broadcast(self.gen_arena.flat, 0)
            torch.distributed.broadcast(self.disc_arena.flat, 0)
        self._vgg_dev = None
        self._loss_buf = None

    # ------------------------------------------------------------------ helpers
    def _vgg_params(self, device):
        conv = self.content_model.features[0]
        ver = (conv.weight._version, conv.bias._version)
        if self._vgg_dev is None or self._vgg_dev[0].device != device or self._vgg_ver != ver:
            self._vgg_ver = ver
            self._vgg_dev = (conv.weight.detach().to(device, torch.float32).contiguous(),
                             conv.bias.detach().to(device, torch.float32).contiguous())
        return self._vgg_dev

    def _vgg_prefix(self, device):
        vp = getattr(self, "_vgg_pre", None)
        if vp is None or vp.device != device or vp.stale():
            from .vgg_prefix import VggPrefix
            vp = self._vgg_pre = VggPrefix(self.content_model, pose_utils.get_layer_ind(self.content_loss_layer), device)
        return vp

    def _loss_readback_begin(self, loss, streams):
        """Queue the device->host copy of the loss scalars on a copy stream right behind the kernels that PRODUCE them,
        not behind the backward pass and the optimiser that follow: the update call then returns as soon as the losses
        exist, and the host enqueues the next update while the GPU is still busy with this one (no idle gap between
        updates).  Returns a token for _loss_readback_end, or None when streams are off / on CPU."""
        from .. import engine as _engine
        dev = loss.device
        if dev.type != "cuda" or not _engine.STREAMS or os.environ.get("PTK_STREAMS", "1") == "0":
            return None
        cs = getattr(self, "_copy_st", None)
        if cs is None or cs.device != dev:
            cs = self._copy_st = torch.cuda.Stream(device=dev)
            self._loss_pin = torch.empty(loss.numel(), dtype=loss.dtype).pin_memory()
        for st in streams:
            if st is not None:
                cs.wait_stream(st)
        with torch.cuda.stream(cs):
            self._loss_pin.copy_(loss, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        loss.record_stream(cs)
        return ev

    def _loss_readback_end(self, token, loss):
        if token is None:
            return loss.tolist()                                         # single device->host sync
        token.synchronize()
        return self._loss_pin.tolist()

    def _opt_stream(self, dev):
        """Stream for the per-bucket all-reduce + Adam (None on CPU / with PTK_STREAMS=0)."""
        from .. import engine as _engine
        if dev.type != "cuda" or not _engine.STREAMS or os.environ.get("PTK_STREAMS", "1") == "0":
            return None
        st = getattr(self, "_ost", None)
        if st is None or st.device != dev:
            st = self._ost = torch.cuda.Stream(device=dev)
        return st

    def _allreduce(self, arena):
        if self.world > 1:
            torch.distributed.all_reduce(arena.grad)

    def _fill_disc_input(self, din, inp, middle, P):
        """din[..., :3+P] = (img, src pose); din[..., 3+P:6+P] = middle (or left to the generator);
        din[..., 6+P:6+2P] = target pose   (pose_gan.py:84-86,131-135)."""
        segs = [(inp, 0, 3 + P, 0), (inp, 3 + P, P, 6 + P)]
        if middle is not None:
            segs.append((middle, 0, 3, 3 + P))
        # one pass writing whole rows (zeros in the generator's slot and in the channel padding)
        K.gather_nhwc(segs, Slice(din, 0, din.shape[-1]), din.shape[-1])

    def _prep(self, t, dtype=torch.float32):
        t = t.cuda() if not t.is_cuda else t
        if t.is_cuda and t.device.index != torch.cuda.current_device():
            # the C ABI launches on the runtime's current device with a raw stream handle: one process drives one GPU
            raise RuntimeError("pose_transfer_b200: tensor on cuda:%d but the current device is cuda:%d -- call "
                               "torch.cuda.set_device() first (one process per GPU)" % (t.device.index, torch.cuda.current_device()))
        return t.to(dtype).contiguous() if t.dtype != dtype else t.contiguous()

    # ------------------------------------------------------------------ updates
    def gen_update(self, input, target, other_inputs, opt, drop=None):
        """pose_gan.py:69-115.  gen_type='stacked' (pose_gan.py:72-77): other_inputs carries interpol_pose [N,S*P,H,W],
        interpol_warps [N,S,10,8], interpol_masks [N,S,10,H,W]; `drop` is then a list of S noise triples (test hook)."""
        if opt['gen_type'] == 'stacked':
            stacked = (self._prep(other_inputs['interpol_pose']), other_inputs['interpol_warps'], other_inputs['interpol_masks'])
            return self._gen_step(input, target, None, None, opt, drop, stacked=stacked)
        return self._gen_step(input, target, other_inputs['warps'], other_inputs['masks'], opt, drop)

    def _forward_gen(self, input, warps, masks, drop, d_input, stacked):
        """Generator forward of an update: (out_gen, outputs_gen list)."""
        if stacked is None:
            out = self.gen.engine.forward(input, warps, masks, drop=drop if drop is not None else self.gen._next_drop(),
                                          repack=None, d_input=d_input)
            return out, []
        pose, swarps, smasks = stacked
        if swarps is not None:
            swarps = swarps.cuda().float() if not swarps.is_cuda else swarps.float()
            smasks = smasks.cuda() if not smasks.is_cuda else smasks
        outs = self.gen.run_stacks(input, pose, swarps, smasks, drops=drop, d_input=d_input, repack=None)
        return outs[-1], outs

    def _gen_step(self, input, target, warps, masks, opt, drop=None, stacked=None):
        P = opt['pose_dim']
        input, target = self._prep(input), self._prep(target)
        N, _, H, W = input.shape
        dev = input.device
        self.gen_arena.check()
        self.disc_arena.check()
        self.gen_arena.zero_grad()                                       # self.gen.zero_grad()  (pose_gan.py:70)
        loss = torch.zeros(4, device=dev)

        din = self.disc.engine.input_buffer(N, H, W, dev)
        self._fill_disc_input(din, input, None, P)
        out_gen, outputs_gen = self._forward_gen(input, warps, masks, drop, Slice(din, 3 + P, 3), stacked)
        # The content loss (VGG conv1_1 + 5x5 NN loss, FFMA-bound) and the adversarial branch (D forward + input
        # gradient, tensor-core / memory bound) only share out_gen: they run concurrently on two streams.
        dpred = self.gen.engine.ws.get("dpred_%d_%d_%d" % (N, H, W), (N, 3, H, W))
        side = self.gen.engine._side_stream(dev)
        main = torch.cuda.current_stream() if side is not None else None
        if side is not None:
            side.wait_stream(main)
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            if self.content_loss_layer != 'none' and pose_utils.get_layer_ind(self.content_loss_layer) == 1:
                # block1_conv2 (the north-star configuration): extractor fused into the loss kernels
                vw, vb = self._vgg_params(dev)
                area = self.nn_loss_area_size
                argmin = self.gen.engine.ws.get("nn_argmin_%d_%d_%d" % (N, H, W), (N, H, W), torch.uint8)
                K.nnloss_forward(out_gen, target, vw, vb, area, opt['l1_penalty_weight'], loss[2:3], argmin)
                K.nnloss_backward(out_gen, target, vw, vb, argmin, area, opt['l1_penalty_weight'], dpred)
            elif self.content_loss_layer != 'none':
                # any other depth: materialised VGG prefix (models/vgg_prefix.py) + nn_loss on its features
                vp = self._vgg_prefix(dev)
                f_tgt = vp.forward(target, "tgt")
                f_gen = vp.forward(out_gen, "gen")
                area = self.nn_loss_area_size
                fN, fC, fH, fW = f_gen.shape
                argmin = vp.ws.get("nn_argmin", (fN, fH, fW), torch.uint8)
                dfeat = vp.ws.get("dfeat", (fN, fC, fH, fW))
                K.nnloss_features_forward(f_gen, f_tgt, area, opt['l1_penalty_weight'], loss[2:3], argmin)
                K.nnloss_features_backward(f_gen, f_tgt, argmin, area, opt['l1_penalty_weight'], dfeat)
                dpred.copy_(vp.backward(dfeat, "gen"))
            else:
                K.l1_loss(out_gen, target, opt['l1_penalty_weight'], loss[2:3], dpred)

        logits = self.disc.engine.forward(din, repack=None)
        J = logits.shape[1]
        dlog4 = self.disc.engine.dlogits_buffer(N, J)
        # ad_loss = sum_n -mean_j log(out+1e-7), * gan_penalty_weight / batch_size   (pose_gan.py:90-98,107)
        K.adv_loss(logits, N, J, N, opt['gan_penalty_weight'] / self.batch_size, loss[0:2], dlog4, dlog4.shape[1])
        readback = self._loss_readback_begin(loss, [torch.cuda.current_stream() if dev.type == "cuda" else None, side])
        din_grad = self.disc.engine.backward(dlog4, grads=None, need_input_grad=True)
        if side is not None:
            main.wait_stream(side)

        # Data parallel: the gradient all-reduce of a sub-network is issued (asynchronously, on NCCL's own stream) as soon
        # as its backward is enqueued, so the 205 MB decoder bucket travels under the encoders' backward passes.
        # ... and that bucket's Adam update follows it on the same "optimiser stream", so on one GPU the 2.3 GB of optimiser
        # traffic also hides under the rest of the backward pass.
        ost = self._opt_stream(dev)
        self.gen_opt.begin_step()
        covered = []

        def stage_done(stage, part=None):
            lo, hi = self.gen_arena.segment_range_of(self.gen.engine.stage_parts(stage, part))
            covered.append((lo, hi))
            if ost is not None:
                ost.wait_stream(torch.cuda.current_stream())
            with (torch.cuda.stream(ost) if ost is not None else contextlib.nullcontext()):
                if self.world > 1:
                    torch.distributed.all_reduce(self.gen_arena.grad[lo:hi])
                self.gen_opt.step_range(lo, hi)

        if stacked is None:
            self.gen.engine.backward(self.gen_arena.grads, dout_nchw=dpred, dout_nhwc=Slice(din_grad, 3 + P, 3),
                                     on_stage=stage_done)
        else:
            self.gen.backward_stacks(self.gen_arena.grads, dpred, Slice(din_grad, 3 + P, 3), on_stage=stage_done)
        assert sorted(covered)[0][0] == 0 and sorted(covered)[-1][1] == self.gen_arena.total and \
            all(a[1] == b[0] for a, b in zip(sorted(covered), sorted(covered)[1:])), "optimiser buckets do not tile the arena"
        if ost is not None:
            torch.cuda.current_stream().wait_stream(ost)
        self.gen_opt.end_step()
        host = self._loss_readback_end(readback, loss)
        self.gen_ad_loss, self.gen_ll_loss = host[0], host[2]
        self.gen_total_loss = float(torch.tensor(host[0]) + torch.tensor(host[2]))
        return out_gen, outputs_gen, [self.gen_total_loss, self.gen_ll_loss, self.gen_ad_loss]

    def dis_update(self, input, target, other_inputs, real_inp, real_target, opt, drop=None):
        """pose_gan.py:117-171 (gen_type='stacked': :120-125, the discriminator sees the LAST stack's output)."""
        if opt['gen_type'] == 'stacked':
            stacked = (self._prep(other_inputs['interpol_pose']), other_inputs['interpol_warps'], other_inputs['interpol_masks'])
            return self._dis_step(input, target, None, None, real_inp, real_target, opt, drop, stacked=stacked)
        return self._dis_step(input, target, other_inputs['warps'], other_inputs['masks'], real_inp, real_target, opt, drop)

    def _dis_step(self, input, target, warps, masks, real_inp, real_target, opt, drop=None, stacked=None):
        P = opt['pose_dim']
        input, real_inp, real_target = self._prep(input), self._prep(real_inp), self._prep(real_target)
        N, _, H, W = input.shape
        dev = input.device
        self.gen_arena.check()
        self.disc_arena.check()
        self.disc_arena.zero_grad()                                      # self.disc.zero_grad() (pose_gan.py:118)
        loss = torch.zeros(4, device=dev)

        M = N + real_inp.shape[0]
        din = self.disc.engine.input_buffer(M, H, W, dev)
        nr = real_inp.shape[0]
        self._fill_disc_input(din[:nr], real_inp, real_target, P)        # real rows first (pose_gan.py:136)
        self._fill_disc_input(din[nr:], input, None, P)
        self._forward_gen(input, warps, masks, drop, Slice(din[nr:], 3 + P, 3), stacked)
        logits = self.disc.engine.forward(din, repack=None)
        J = logits.shape[1]
        dlog4 = self.disc.engine.dlogits_buffer(M, J)
        # rows < opt['batch_size'] are "true", the rest "fake"; both * gan_w / self.batch_size (pose_gan.py:140-163)
        K.adv_loss(logits, M, J, opt['batch_size'], opt['gan_penalty_weight'] / self.batch_size, loss[0:2], dlog4, dlog4.shape[1])
        readback = self._loss_readback_begin(loss, [torch.cuda.current_stream() if dev.type == "cuda" else None])
        self.disc.engine.backward(dlog4, grads=self.disc_arena.grads, need_input_grad=False)
        # all-reduce + Adam of the discriminator on the optimiser stream: they run while the host reads the losses back
        # and sets up the next update; the compute stream re-joins (device-side dependency, no host sync)
        ost = self._opt_stream(dev)
        if ost is not None:
            ost.wait_stream(torch.cuda.current_stream())
        with (torch.cuda.stream(ost) if ost is not None else contextlib.nullcontext()):
            self._allreduce(self.disc_arena)
            self.disc_opt.step()
        host = self._loss_readback_end(readback, loss)
        if ost is not None:
            torch.cuda.current_stream().wait_stream(ost)
        self.dis_true_loss, self.dis_fake_loss = host[0], host[1]
        self.dis_total_loss = float(torch.tensor(host[0]) + torch.tensor(host[1]))
        return [self.dis_total_loss, self.dis_true_loss, self.dis_fake_loss]

    def nn_loss(self, predicted, ground_truth, nh=3, nw=3):
        """pose_gan.py:173-199 on materialised feature tensors [N,C,H,W] (differentiable w.r.t. `predicted`).  The training
        step itself uses ptk_nnloss_forward/backward, which fuse the VGG feature extractor into this loss."""
        if nh != nw:
            # the reference's ConstantPad2d((v_pad, v_pad, h_pad, h_pad)) mixes the axes: only square windows run there
            raise RuntimeError("nn_loss: the reference only supports square windows (nh == nw)")
        return _NNLossFn.apply(predicted, ground_truth, int(nh))

    # ------------------------------------------------------------------ checkpoints (pose_gan.py:201-220)
    def resume(self, save_dir):
        last_model_name = pose_utils.get_model_list(save_dir, "gen")
        if last_model_name is None:
            return 1
        self.gen.load_state_dict(torch.load(last_model_name))
        epoch = int(last_model_name[-7:-4])
        print('Resume gen from epoch %d' % epoch)
        last_model_name = pose_utils.get_model_list(save_dir, "dis")
        if last_model_name is None:
            return 1
        epoch = int(last_model_name[-7:-4])
        self.disc.load_state_dict(torch.load(last_model_name))
        print('Resume disc from epoch %d' % epoch)
        return epoch

    def save(self, save_dir, epoch):
        if self.rank != 0:
            return
        gen_filename = os.path.join(save_dir, 'gen_{0:03d}.pkl'.format(epoch))
        disc_filename = os.path.join(save_dir, 'disc_{0:03d}.pkl'.format(epoch))
        torch.save(self.gen.state_dict(), gen_filename)
        torch.save(self.disc.state_dict(), disc_filename)

    def normalize_image(self, x):
        return x[:, 0:3, :, :]


class Pose_GAN(DeformablePose_GAN):
    """Drop-in for src_baseline/models/pose_gan.py:10-170: single-encoder Generator, PatchGAN, adversarial +
    L1 loss, xavier initialisation, Adam(2e-4, (0.5, 0.999)); update methods take `interpol_pose` where the deformable
    trainer takes `other_inputs` (unused for gen_type='baseline').  Runs on the same engines / kernels."""

    def __init__(self, opt):
        nn.Module.__init__(self)
        check_mode = getattr(opt, "checkMode", 0)
        if check_mode == 0:
            nfilters_decoder = (512, 512, 512, 256, 128, 3) if max(opt.image_size) < 256 else (512, 512, 512, 512, 256, 128, 3)
            nfilters_encoder = (64, 128, 256, 512, 512, 512) if max(opt.image_size) < 256 else (64, 128, 256, 512, 512, 512, 512)
        else:       # --checkMode: reduced nets for over-fitting checks (src_baseline/models/pose_gan.py:16-21)
            nfilters_decoder = (128, 3) if max(opt.image_size) < 256 else (256, 128, 3)
            nfilters_encoder = (64, 128) if max(opt.image_size) < 256 else (64, 128, 256)
        input_nc = 3 + 2 * opt.pose_dim if opt.use_input_pose else 3 + opt.pose_dim
        if not opt.use_input_pose:
            raise NotImplementedError("only use_input_pose=True is on the B200 path")
        self.num_stacks = opt.num_stacks
        self.batch_size = opt.batch_size
        self.pose_dim = opt.pose_dim
        self.image_size = tuple(opt.image_size)
        if opt.gen_type == 'stacked':
            self.gen = Stacked_Baseline_Generator(input_nc, opt.num_stacks, opt.pose_dim, nfilters_encoder, nfilters_decoder,
                                                  use_input_pose=opt.use_input_pose)
        elif opt.gen_type == 'baseline':
            self.gen = Generator(input_nc, nfilters_encoder, nfilters_decoder, use_input_pose=opt.use_input_pose)
        else:
            raise Exception('Invalid gen_type')
        self.disc = Discriminator(input_nc + 3, use_input_pose=opt.use_input_pose, checkMode=check_mode, baseline_tree=True)
        self.disc.apply(xavier_weights_init)            # src_baseline/models/pose_gan.py:51-52
        self.gen.apply(xavier_weights_init)
        base = argparse_like(opt, content_loss_layer='none', nn_loss_area_size=1)
        self._finish_init(base)

    def gen_update(self, input, target, interpol_pose, opt, drop=None):
        stacked = (self._prep(interpol_pose), None, None) if opt['gen_type'] == 'stacked' else None
        return self._gen_step(input, target, None, None, opt, drop, stacked=stacked)

    def dis_update(self, input, target, interpol_pose, real_inp, real_target, opt, drop=None):
        stacked = (self._prep(interpol_pose), None, None) if opt['gen_type'] == 'stacked' else None
        return self._dis_step(input, target, None, None, real_inp, real_target, opt, drop, stacked=stacked)


def argparse_like(opt, **overrides):
    """A shallow copy of an options namespace with some attributes replaced."""
    import argparse
    d = dict(vars(opt))
    d.update(overrides)
    return argparse.Namespace(**d)
