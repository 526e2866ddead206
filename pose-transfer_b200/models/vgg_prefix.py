"""VGG-19 feature prefix for content_loss_layer deeper than block1_conv2.

The reference's Feature_Extractor (utils/pose_utils.py:320-338) runs `model.features[0..layer]` of torchvision's VGG-19
on the view-normalised image; models/pose_gan.py:100-104 back-propagates the 5x5 nearest-neighbour L1 loss through it
into the generator output.  For 'block1_conv2' (layer index 1) the trainer uses the fused ptk_nnloss_* kernels; every
other depth runs here: 3x3 convolutions through the same tcgen05 implicit-GEMM kernels as the networks (ConvLayer),
2x2 max-pooling, ReLU and the pre-processing as small streaming kernels.  The VGG weights are frozen: only the input
gradient (dgrad) is computed."""
import torch
import torch.nn as nn

from .. import kernels as K
from ..engine import ConvLayer, _Workspace, SPLITK_SCRATCH
from ..kernels import Slice, ACT_NONE, ACT_RELU


class VggPrefix:
    def __init__(self, model, layer_ind, device):
        mods = list(model.features.children())
        if not 0 <= layer_ind < len(mods):
            raise ValueError("content_loss_layer index %d outside vgg19.features (0..%d)" % (layer_ind, len(mods) - 1))
        mods = mods[:layer_ind + 1]
        self.device = device
        self.ops = []                      # ("conv", ConvLayer, fused_relu) | ("pool",)
        self._versions = []
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d):
                if m.kernel_size != (3, 3) or m.stride != (1, 1) or m.padding != (1, 1):
                    raise NotImplementedError("VggPrefix: only 3x3 / stride 1 / pad 1 convolutions (torchvision vgg19.features)")
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                w = m.weight.detach().to(device, torch.float32).contiguous()
                b = m.bias.detach().to(device, torch.float32).contiguous()
                layer = ConvLayer(w, b, False, 3, 1, 1)
                layer.pack_forward()
                self.ops.append(("conv", layer, relu))
                self._versions += [(m.weight, m.weight._version), (m.bias, m.bias._version)]
                i += 2 if relu else 1
            elif isinstance(m, nn.MaxPool2d):
                ks = m.kernel_size if isinstance(m.kernel_size, tuple) else (m.kernel_size, m.kernel_size)
                if ks != (2, 2):
                    raise NotImplementedError("VggPrefix: only MaxPool2d(2, 2)")
                self.ops.append(("pool",))
                i += 1
            else:                                 # (a ReLU always follows a conv in vgg19.features and is fused above)
                raise NotImplementedError("VggPrefix: unsupported module %s" % type(m).__name__)
        self.ws = _Workspace(device)
        self.saved = {}

    def stale(self):
        """True if the content model's weights were modified in place since they were packed."""
        return any(w._version != v for w, v in self._versions)

    def out_shape(self, N, H, W):
        C = 3
        for op in self.ops:
            if op[0] == "conv":
                C = op[1].cout
            elif op[0] == "pool":
                H, W = H // 2, W // 2
        return N, C, H, W

    def forward(self, x, tag):
        """x NCHW [N,3,H,W] -> features NCHW; the activations are kept under `tag` for backward(tag)."""
        N, C, H, W = x.shape
        assert C == 3
        ws = self.ws
        scratch = ws.get("splitk", (SPLITK_SCRATCH,))
        first = self.ops[0][1] if self.ops[0][0] == "conv" else None
        cin_pad = first.cin_pad if first is not None else 4
        cur = ws.get("%s_in" % tag, (N, H, W, cin_pad))        # zero-initialised once; channels 3.. are never written
        K.vgg_preprocess(x.contiguous(), cur)
        acts = [(cur, H, W, 3)]
        h, w, c = H, W, 3
        for i, op in enumerate(self.ops):
            if op[0] == "conv":
                layer = op[1]
                y = ws.get("%s_a%d" % (tag, i), (N, h, w, layer.cout))
                layer.forward(Slice(cur), N, h, w, Slice(y), act=ACT_RELU if op[2] else ACT_NONE, scratch=scratch)
                c = layer.cout
            else:
                y = ws.get("%s_a%d" % (tag, i), (N, h // 2, w // 2, c))
                K.maxpool2_forward(cur, y, N, h, w, c)
                h, w = h // 2, w // 2
            cur = y
            acts.append((cur, h, w, c))
        out = ws.get("%s_feat" % tag, (N, c, h, w))
        K.nhwc_to_nchw(Slice(cur), out)
        self.saved[tag] = acts
        return out

    def backward(self, dfeat, tag):
        """d loss / d features (NCHW) -> d loss / d x (NCHW [N,3,H,W]) for the forward pass recorded under `tag`."""
        acts = self.saved[tag]
        N = dfeat.shape[0]
        ws = self.ws
        scratch = ws.get("splitk", (SPLITK_SCRATCH,))
        _, h, w, c = acts[-1]
        g = ws.get("g%d" % len(self.ops), (N, h, w, c))
        K.nchw_to_nhwc(dfeat.contiguous(), 0, c, Slice(g))
        for i in range(len(self.ops) - 1, -1, -1):
            op = self.ops[i]
            x_in, hi, wi, ci = acts[i]
            y_out = acts[i + 1][0]
            if op[0] == "conv":
                layer = op[1]
                if op[2]:
                    K.relu_backward(y_out, g, N * h * w, c)
                first = i == 0
                dx = ws.get("g%d" % i, (N, hi, wi, layer.cin_pad if first else layer.cin))
                layer.dgrad(Slice(g), N, hi, wi, Slice(dx), dx_channels=layer.cin_pad if first else None, scratch=scratch)
            else:
                if hi % 2 or wi % 2:
                    raise NotImplementedError("VggPrefix.backward: max-pool over an odd extent (%dx%d)" % (hi, wi))
                dx = ws.get("g%d" % i, (N, hi, wi, ci))
                K.maxpool2_backward(g, x_in, dx, N, hi, wi, ci)
            g, h, w, c = dx, hi, wi, ci
        dpred = ws.get("dpred", (N, 3, h, w))
        K.vgg_preprocess_backward(g, dpred)
        return dpred
