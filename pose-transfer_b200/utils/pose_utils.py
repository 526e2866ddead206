"""Drop-in for the hot-path part of the reference's ``utils/pose_utils.py``
(src_deformable/utils/pose_utils.py:45-54, 227-233, 312-338).  Display / drawing / key-point helpers of
the reference file are host-side visualisation code and stay out of scope."""
import os

import torch

from .. import kernels as K

MISSING_VALUE = -1


def get_model_list(dirname, key):
    """utils/pose_utils.py:45-54: lexicographically last checkpoint containing `key` and 'pkl'."""
    if os.path.exists(dirname) is False:
        return None
    gen_models = [os.path.join(dirname, f) for f in os.listdir(dirname)
                  if os.path.isfile(os.path.join(dirname, f)) and key in f and "pkl" in f]
    if gen_models is None or gen_models == []:
        return None
    gen_models.sort()
    return gen_models[-1]


def get_imgpose(input, use_input_pose, pose_dim):
    """utils/pose_utils.py:227-233 (views only)."""
    inp_img = input[:, :3]
    inp_pose = input[:, 3:3 + pose_dim] if use_input_pose else None
    tg_pose_index = 3 + pose_dim if use_input_pose else 6
    tg_pose = input[:, tg_pose_index:]
    return inp_img, inp_pose, tg_pose


def get_layer_ind(layer_name):
    """utils/pose_utils.py:312-317 ('block1_conv2' -> 1, i.e. features[0..1] = conv1_1 + ReLU)."""
    block, conv = layer_name.split('_')
    block = int(block[-1])
    conv = int(conv[-1])
    blocks = ['0', '5', '10', '19', '28']
    return int(blocks[block - 1]) + conv - 1


VGG_MEAN = (0.485, 0.456, 0.406)
VGG_STD = (0.229, 0.224, 0.225)


def Feature_Extractor(model, input=None, layer_name=None):
    """utils/pose_utils.py:320-338: model.features[0..layer] on the view-normalised image, NCHW in / NCHW out.
    Layer index 1 ('block1_conv2' = relu(conv1_1), the north-star depth) is one ptk conv; deeper layers run the
    VggPrefix (models/vgg_prefix.py).  (The training step at 'block1_conv2' never calls this: ptk_nnloss_* fuse the
    extractor into the loss.)"""
    layer = get_layer_ind(layer_name)
    if not input.is_cuda:
        raise RuntimeError("Feature_Extractor: CUDA tensors required (no CPU fallback)")
    if layer != 1:
        from ..models.vgg_prefix import VggPrefix
        return VggPrefix(model, layer, input.device).forward(input.float(), "fx").clone()
    conv = model.features[0]
    w = conv.weight.detach().to(input.device, torch.float32).contiguous()
    b = conv.bias.detach().to(input.device, torch.float32).contiguous()
    N, C, H, W = input.shape
    # preprocessing: element with flat per-sample index i uses mean[i % 3], std[i % 3]  (pose_utils.py:324-331)
    idx = torch.arange(C * H * W, device=input.device) % 3
    mean = torch.tensor(VGG_MEAN, device=input.device)[idx].view(1, C, H, W)
    std = torch.tensor(VGG_STD, device=input.device)[idx].view(1, C, H, W)
    x = ((input - mean) / std).contiguous()
    xin = torch.zeros(N, H, W, 4, device=input.device)
    K.nchw_to_nhwc(x, 0, 3, K.Slice(xin, 0, 3))
    wt = torch.empty(9 * 4 * 64, device=input.device)
    K.pack_weight(w, wt, 64, 3, 9, 64, 4, 1)
    y = torch.empty(N, H, W, 64, device=input.device)
    g = K.conv_geom(N, H, W, 4, 4, H, W, 64, 64, 3, 1, 1, False, K.IMPL_SIMT)
    K.conv_forward(g, xin, wt, None, b, K.ACT_RELU, y)
    out = torch.empty(N, 64, H, W, device=input.device)
    K.nhwc_to_nchw(K.Slice(y), out)
    return out
