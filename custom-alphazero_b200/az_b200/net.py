"""Policy/value network of the reference, restated for PyTorch (the only dense contraction on the
self-play path).  PolicyValueNet is the trainable fp32 module; InferenceNet folds it for inference and routes the
forward pass: on the 6x7 headline shapes the WHOLE net is one hand-written tcgen05 kernel (az_net_forward, csrc/az_tower.cu),
for 8x8 / chess the residual tower is that kernel without its two ends (az_net_tower) between hand-written stem / head
kernels, and cuDNN's fused-epilogue convolutions remain as the comparison arm and for boards the kernel's tiling does not
cover (9x9).

Architecture follows /root/reference/custom_alphazero/model/tensorflow/model.py:21-188 and
base_layers.py:20-125 (the TensorFlow/Keras model; the PyTorch copy in the reference is dead code):
  input  x [B, H, W, 4] float (Board.full_state, NHWC)
  stem   Conv3x3(4 -> F) + BN + ReLU                                     (model.py:36-46)
  tower  depth x { Conv3x3+BN+ReLU -> Conv3x3+BN ; shortcut Conv1x1+BN of the block INPUT ;
                   add ; ReLU }                                          (base_layers.py:85-125)
  policy Conv1x1(F -> 2)+BN+ReLU -> Flatten (NHWC order) -> Dense(A) softmax     (model.py:68-103)
  value  Conv1x1(F -> 1)+BN+ReLU -> Flatten -> Dense(256) ReLU -> Dense(1) tanh  (model.py:106-149)
Keras defaults: Conv2D/Dense use_bias=True, glorot-uniform kernels, zero biases; BatchNormalization
eps 1e-3, momentum 0.99.  F = 128, depth = 4 (config.py:63,71): 1 267 037 trainable parameters at 6x7.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-3
BN_MOMENTUM = 0.01  # Keras momentum 0.99


def flops_per_eval(height, width, n_actions, filters=128, depth=4, in_planes=4):
    """2 * MAC over the layers above: 105 037 976 at 6x7 / A=7."""
    px = height * width
    macs = px * 9 * in_planes * filters
    macs += depth * (2 * px * 9 * filters * filters + px * filters * filters)
    macs += px * filters * 2 + 2 * px * n_actions
    macs += px * filters * 1 + px * 256 + 256
    return 2 * macs


def _split_stage_halves(img):
    """Stage images [.., 8 chunks, 128 cout, 8] -> the CTA-pair layout [.., 2 halves, 8 chunks, 64 cout, 8] (flat)."""
    t = img.reshape(-1, 8, 2, 64, 8)                 # stage, chunk, half, cout, e
    return t.permute(0, 2, 1, 3, 4).reshape(-1)      # stage, half, chunk, cout, e


def pack_tower_weights(blocks, pair=False):
    """[(w1, b1, w2, wp, b2p)] per block (BN folded; w [128, 128, k, k], any float dtype) -> (w_img bf16 [depth * 38 * 8192],
    bias float32 [depth, 2, 128]) in the layout az_net_tower streams (include/az_b200.h): a 16 KB stage is one filter tap
    x 64 input channels as the K-major, unswizzled UMMA B operand [8 chunks][128 cout][8 cin]; stages come in
    consumption order: conv1 = 2 input-channel halves x taps (ky, kx) row-major, the 1x1 shortcut x 2 halves, conv2 like
    conv1 (channel-half major, so the MMAs on channels 0-63 can start while the epilogue still writes 64-127)."""
    imgs, biases = [], []
    for w1, b1, w2, wp, b2p in blocks:
        for w in (w1, wp, w2):
            cout, cin, kh, kw = w.shape
            assert cout == 128 and cin == 128, "az_net_tower is built for 128 filters"
            t = w.detach().float().permute(2, 3, 1, 0).reshape(kh * kw, 2, 8, 8, cout)  # tap, half, chunk, e, cout
            imgs.append(t.permute(1, 0, 2, 4, 3).reshape(-1))                             # half, tap, chunk, cout, e
        biases.append(torch.stack([b1.detach().float(), b2p.detach().float()]))
    img = torch.cat(imgs)
    if pair:  # layout 1 of az_net_tower: CTA r of a pair streams output channels 64 r .. 64 r + 63 of every stage
        img = _split_stage_halves(img)
    return img.to(torch.bfloat16).contiguous(), torch.stack(biases).contiguous()


def pack_stem_weights(w, pair=False):
    """Folded stem weights [128, 4, 3, 3] -> the 3 stages (bf16, 3 * 8192 elements) az_net_forward reads before the tower
    image: tap t = (ky, kx) row-major sits in stage t // 4 at chunk columns 2 * (t % 4) (planes in K 0-3, zeros in K 4-7)
    and 2 * (t % 4) + 1 (zeros): each tap is one K = 16 MMA against chunk columns 0-1 of the activation buffers."""
    cout, cin, kh, kw = w.shape
    assert cout == 128 and cin == 4 and kh == 3 and kw == 3
    img = torch.zeros((3, 8, cout, 8), dtype=torch.float32)
    wt = w.detach().float().cpu().permute(2, 3, 0, 1).reshape(9, cout, cin)  # tap, cout, plane
    for t in range(9):
        img[t // 4, 2 * (t % 4), :, :4] = wt[t]
    flat = img.reshape(-1)
    if pair:
        flat = _split_stage_halves(flat)
    return flat.to(torch.bfloat16).contiguous()


class ConvBN(nn.Module):
    def __init__(self, cin, cout, k, relu):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=k // 2, bias=True)
        self.bn = nn.BatchNorm2d(cout, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.relu = relu
        nn.init.xavier_uniform_(self.conv.weight)
        nn.init.zeros_(self.conv.bias)

    def forward(self, x):
        y = self.bn(self.conv(x))
        return F.relu(y) if self.relu else y

    def folded(self):
        """(weight, bias) of the equivalent conv with the BN inference transform folded in."""
        s = self.bn.weight / torch.sqrt(self.bn.running_var + self.bn.eps)
        w = self.conv.weight * s[:, None, None, None]
        b = (self.conv.bias - self.bn.running_mean) * s + self.bn.bias
        return w, b


class ResBlock(nn.Module):
    def __init__(self, filters):
        super().__init__()
        self.c1 = ConvBN(filters, filters, 3, relu=True)
        self.c2 = ConvBN(filters, filters, 3, relu=False)
        self.proj = ConvBN(filters, filters, 1, relu=False)

    def forward(self, x):
        return F.relu(self.proj(x) + self.c2(self.c1(x)))


class PolicyValueNet(nn.Module):
    """Trainable fp32 module (the checker for the bf16 inference path and the thing a trainer updates)."""

    def __init__(self, height=6, width=7, n_actions=7, filters=128, depth=4, in_planes=4):
        super().__init__()
        self.height, self.width, self.n_actions, self.filters, self.depth = height, width, n_actions, filters, depth
        self.in_planes = in_planes  # 4 for Connect-N (board.py:83-98), 118 for chess (chess/board.py:58-73)
        self.stem = ConvBN(in_planes, filters, 3, relu=True)
        self.blocks = nn.ModuleList([ResBlock(filters) for _ in range(depth)])
        self.policy_conv = ConvBN(filters, 2, 1, relu=True)
        self.policy_fc = nn.Linear(2 * height * width, n_actions)
        self.value_conv = ConvBN(filters, 1, 1, relu=True)
        self.value_fc1 = nn.Linear(height * width, 256)
        self.value_fc2 = nn.Linear(256, 1)
        for fc in (self.policy_fc, self.value_fc1, self.value_fc2):
            nn.init.xavier_uniform_(fc.weight)
            nn.init.zeros_(fc.bias)

    def trunk(self, x_nhwc):
        x = x_nhwc.permute(0, 3, 1, 2)  # logical NCHW over NHWC memory
        x = self.stem(x)
        for b in self.blocks:
            x = b(x)
        return x

    def heads(self, x):
        B = x.shape[0]
        p = self.policy_conv(x).permute(0, 2, 3, 1).reshape(B, -1)  # Keras Flatten on NHWC
        v = self.value_conv(x).permute(0, 2, 3, 1).reshape(B, -1)
        logits = self.policy_fc(p)
        value = torch.tanh(self.value_fc2(F.relu(self.value_fc1(v))))
        return logits, value

    def forward(self, x_nhwc):
        """-> (policy [B, A] softmax, value [B, 1] tanh), like PolicyValueModel.call (model.py:182-188)."""
        logits, value = self.heads(self.trunk(x_nhwc))
        return torch.softmax(logits, dim=-1), value

    def n_parameters(self):
        return sum(p.numel() for p in self.parameters())

    # ---- the reference's own weight format: the list PolicyValueModel.get_weights() returns (model.py:168-176 hashes it,
    # save_weights / load_weights at :190-212 store it).  Order as Keras 2.7 builds it (poetry.lock): per top-level layer
    # (residual_tower, policy_head, value_head) all trainable variables in layer order, then the BatchNormalization
    # moving statistics; kernels HWIO / [in, out].  A maintainer of the reference exports a checkpoint with
    # `np.savez(path, *model.get_weights())` and loads it here with `net.load_keras_weights(list(np.load(path).values()))`.
    def _keras_groups(self):
        tower = [self.stem] + [cb for blk in self.blocks for cb in (blk.c1, blk.c2, blk.proj)]
        return ((tower, []), ([self.policy_conv], [self.policy_fc]), ([self.value_conv], [self.value_fc1, self.value_fc2]))

    def to_keras_weights(self):
        """-> list of numpy arrays in PolicyValueModel.get_weights() order."""
        out = []
        for convs, denses in self._keras_groups():
            for cb in convs:
                out += [cb.conv.weight.detach().permute(2, 3, 1, 0), cb.conv.bias.detach(), cb.bn.weight.detach(), cb.bn.bias.detach()]
            for fc in denses:
                out += [fc.weight.detach().t(), fc.bias.detach()]
            for cb in convs:
                out += [cb.bn.running_mean, cb.bn.running_var]
        return [t.cpu().contiguous().numpy().copy() for t in out]

    def load_keras_weights(self, weights):
        """Inverse of to_keras_weights: fills this module from a PolicyValueModel.get_weights() list (shapes checked)."""
        weights = list(weights)
        it = iter(weights)

        def put(dst, arr, what):
            src = torch.as_tensor(np.asarray(arr), dtype=dst.dtype)
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"load_keras_weights: {what} has shape {tuple(src.shape)}, expected {tuple(dst.shape)}")
            with torch.no_grad():
                dst.copy_(src)

        n_expected = sum(6 * len(c) + 2 * len(d) for c, d in self._keras_groups())
        if len(weights) != n_expected:
            raise ValueError(f"load_keras_weights: {len(weights)} arrays, this architecture has {n_expected}")
        for convs, denses in self._keras_groups():
            for cb in convs:
                put(cb.conv.weight, np.transpose(np.asarray(next(it)), (3, 2, 0, 1)), "a convolution kernel")
                put(cb.conv.bias, next(it), "a convolution bias")
                put(cb.bn.weight, next(it), "a BatchNormalization gamma")
                put(cb.bn.bias, next(it), "a BatchNormalization beta")
            for fc in denses:
                put(fc.weight, np.asarray(next(it)).T, "a dense kernel")
                put(fc.bias, next(it), "a dense bias")
            for cb in convs:
                put(cb.bn.running_mean, next(it), "a BatchNormalization moving mean")
                put(cb.bn.running_var, next(it), "a BatchNormalization moving variance")
        return self


class InferenceNet(nn.Module):
    """Inference-only copy: BN folded into the convolutions, channels-last, one dtype (bf16 on the
    GPU path, fp32 for the CPU checker).  Outputs float32 (policy [B, A], value [B])."""

    def __init__(self, net: PolicyValueNet, dtype=torch.bfloat16, device="cuda"):
        super().__init__()
        net = net.eval()
        self.height, self.width, self.n_actions = net.height, net.width, net.n_actions
        self.dtype = dtype
        cl = torch.channels_last

        def conv_params(cb):
            w, b = cb.folded()
            return (nn.Parameter(w.detach().to(device=device, dtype=dtype).contiguous(memory_format=cl), requires_grad=False),
                    nn.Parameter(b.detach().to(device=device, dtype=dtype), requires_grad=False))

        self.stem_w, self.stem_b = conv_params(net.stem)
        # cuDNN's tensor-core kernels want the channel count a multiple of 8 (bf16): the chess input (118 planes) is
        # padded to 120 with zero planes against zero weights
        self.in_pad = (-net.in_planes) % 8
        if self.in_pad:
            wpad = F.pad(self.stem_w.data.contiguous(), (0, 0, 0, 0, 0, self.in_pad)).contiguous(memory_format=cl)
            self.stem_w_pad = nn.Parameter(wpad, requires_grad=False)
        self.block_params = nn.ParameterList()
        for blk in net.blocks:
            w1, b1 = conv_params(blk.c1)
            w2, b2 = conv_params(blk.c2)
            wp, bp = conv_params(blk.proj)
            # the two biases that meet at the add are summed once here
            b2p = nn.Parameter((b2.float() + bp.float()).to(dtype), requires_grad=False)
            self.block_params.extend([w1, b1, w2, wp, b2p])
        self.depth = len(net.blocks)
        # the whole tower as one persistent tcgen05 kernel (az_net_tower, csrc/az_tower.cu) where its tiling fits:
        # 128 filters, whole positions in a 128-row tile with (positions * W + 1) <= 22 padding rows (6x7, 8x8, ...)
        import os as _os0

        cells_ = net.height * net.width
        self.fused_tower = (torch.device(device).type == "cuda" and dtype == torch.bfloat16 and net.filters == 128
                            and 1 <= self.depth <= 6 and cells_ <= 128 and (128 // cells_) * net.width + 1 <= 22
                            and (128 // cells_) * cells_ >= 96 and _os0.environ.get("AZ_FUSED_TOWER", "1") != "0")
        # CTA pairs (cta_group::2) halve the weight traffic through each SM's shared memory; AZ_TOWER_PAIR=0 = one CTA per tile
        self.tower_layout = 1 if _os0.environ.get("AZ_TOWER_PAIR", "1") != "0" else 0
        if self.fused_tower:
            img, tb = pack_tower_weights([tuple(self.block_params[5 * i: 5 * i + 5]) for i in range(self.depth)],
                                         pair=bool(self.tower_layout))
            self.tower_img = nn.Parameter(img.to(device), requires_grad=False)
            self.tower_bias = nn.Parameter(tb.to(device), requires_grad=False)
        # ... and the whole net as one kernel (az_net_forward: stem + tower + both heads) for 4-plane boards of up to
        # 48 cells with a narrow action space (Connect-N)
        ppt_ = 128 // cells_
        self.fused_net = (self.fused_tower and net.in_planes == 4 and cells_ <= 48 and ppt_ * net.n_actions <= 32
                          and net.value_fc1.out_features == 256 and _os0.environ.get("AZ_FUSED_NET", "1") != "0")
        if self.fused_net:
            self.net_img = nn.Parameter(torch.cat([pack_stem_weights(self.stem_w.float(), pair=bool(self.tower_layout)).to(device), self.tower_img.data]),
                                        requires_grad=False)
            del self.tower_img  # one copy: the tower image is the tail of the net image (a view, not a second parameter)
            self.tower_img = self.net_img.data[3 * 8192:]
        f32 = lambda t: nn.Parameter(t.detach().to(device=device, dtype=torch.float32).contiguous(), requires_grad=False)  # noqa: E731
        # float32 copies for the hand-written stem / heads kernels (az_net_stem, az_net_heads)
        sw, sb = net.stem.folded()
        self.stem_w32, self.stem_b32 = f32(sw), f32(sb)
        if net.in_planes == 4:
            # the stem as a [128][64] bf16 GEMM operand, K = tap * 4 + plane, for az_net_stem_tc (tcgen05)
            k_major = sw.detach().permute(0, 2, 3, 1).reshape(sw.shape[0], 36)
            self.stem_w16_k = nn.Parameter(F.pad(k_major, (0, 28)).to(device=device, dtype=torch.bfloat16).contiguous(),
                                           requires_grad=False)
        import os as _os

        self.split_heads = _os.environ.get("AZ_SPLIT_HEADS", "0") == "1"
        self.tc_stem = (net.in_planes == 4 and net.filters == 128 and net.height * net.width <= 128
                        and _os.environ.get("AZ_TC_STEM", "1") != "0")
        # both 1x1 head convolutions as one [3, C] matrix (2 policy planes + 1 value plane)
        pw, pb = net.policy_conv.folded()
        vw, vb = net.value_conv.folded()
        self.head_w32 = f32(torch.cat([pw, vw], 0).reshape(3, -1))
        self.head_b32 = f32(torch.cat([pb, vb], 0))
        self.pfc_w, self.pfc_b = f32(net.policy_fc.weight), f32(net.policy_fc.bias)
        if net.n_actions > 128:
            # wide policy layer (chess): bf16 operands for the tensor cores, rows padded to a multiple of 128 for
            # az_net_dense_heads; the value hidden layer likewise
            w16 = net.policy_fc.weight.detach().to(device=device, dtype=torch.bfloat16)
            self.pfc_w16 = nn.Parameter(F.pad(w16, (0, 0, 0, (-net.n_actions) % 128)).contiguous(), requires_grad=False)
            self.v1_w16 = nn.Parameter(net.value_fc1.weight.detach().to(device=device, dtype=torch.bfloat16).contiguous(),
                                       requires_grad=False)
        self.fused_dense_heads = True  # az_net_dense_heads where it applies (GPU, 64 cells, wide policy); else cuBLAS
        if net.in_planes == 118 and net.height == 8 and net.width == 8 and net.filters == 128:
            # az_chess_stem: the stem restricted to the planes that vary on the self-play path (current entry 98-111 and
            # scalars 112-117, padded to 24) + the per-cell constant (bias + the initial position in history entry 6:
            # planes 84-97), from the same bf16-rounded weights the cuDNN stem uses
            w16 = self.stem_w.detach().float()  # [128, 118, 3, 3]
            red = torch.zeros((net.filters, 24, 3, 3), dtype=torch.float32, device=w16.device)
            red[:, :14] = w16[:, 98:112]
            red[:, 14:20] = w16[:, 112:118]
            self.chess_stem_w = nn.Parameter(red.reshape(net.filters, 24, 9).contiguous(), requires_grad=False)
            # the same weights as a [128][256] bf16 GEMM operand, K = tap * 24 + plane (az_chess_stem_tc)
            k_major = red.reshape(net.filters, 24, 9).permute(0, 2, 1).reshape(net.filters, 216)
            self.chess_stem_w16 = nn.Parameter(F.pad(k_major, (0, 40)).to(torch.bfloat16).contiguous(), requires_grad=False)
            # the stem restricted to planes 84-117 (initial-position entry, current entry, scalars; 34 planes padded to
            # 40): on the self-play path the six older history entries are always empty (az_chess_step plane_first = 84)
            tail = F.pad(self.stem_w.detach().contiguous()[:, 84:118], (0, 0, 0, 0, 0, 6)).contiguous(memory_format=cl)
            self.stem_w_tail = nn.Parameter(tail, requires_grad=False)
            from .chess import position_from_fen, unpack_position

            arr = torch.as_tensor(unpack_position(position_from_fen())["array"].astype("int64"))
            start = torch.zeros((1, 14, 8, 8), dtype=torch.float32, device=w16.device)
            start[0, :13] = F.one_hot(arr % 13, 13).permute(2, 0, 1).float()  # np.eye(13)[array] with negative wrap
            cmap = F.conv2d(start, w16[:, 84:98], self.stem_b.detach().float(), padding=1)  # [1, 128, 8, 8]
            # [2][64][128]: map 0 for boards that went through play() + mirror() (initial position in history entry 6), map 1
            # = the bias alone for the un-mirrored ply-0 root, whose deque is seven empty entries (chess/board.py:37-40)
            fresh = self.stem_b.detach().float()[None, :].expand(64, net.filters)
            self.chess_stem_map = nn.Parameter(torch.stack([cmap[0].permute(1, 2, 0).reshape(64, net.filters), fresh]).contiguous(),
                                               requires_grad=False)
        self.v1_w, self.v1_b = f32(net.value_fc1.weight), f32(net.value_fc1.bias)
        # az_net_heads wants the policy rows padded to an odd stride and the value weights transposed
        # (bank-conflict-free shared memory images that the kernel copies verbatim)
        cells = net.height * net.width
        self.pfc_w_pad = f32(F.pad(net.policy_fc.weight.detach(), (0, 1)))
        self.v1_w_pad = f32(net.value_fc1.weight.detach().t())  # transposed [cells][256] for 128-bit smem loads
        self.v2_w, self.v2_b = f32(net.value_fc2.weight), f32(net.value_fc2.bias)
        self.filters = net.filters
        self._net_heads = None
        dev = torch.device(device)
        # fast path: custom stem/heads kernels + cuDNN fused-epilogue tower (GPU, bf16, 128 filters)
        self.fast = (dev.type == "cuda" and dtype == torch.bfloat16 and net.filters == 128 and net.value_fc1.out_features == 256
                     and net.in_planes == 4 and net.n_actions <= 128)
        self._head_struct = None
        self.overlap_shortcut = False
        # the 1x1 projection shortcut through the hand-written tcgen05 GEMM (az_net_conv1x1) instead of cuDNN: both sit
        # on the memory roofline (14.2 vs 12.9 us at 172 032 cells, 6.2 vs 6.8 TB/s of algorithmic traffic with the input
        # still in L2); cuDNN's is 2 us faster inside the tower, so the hand-written one is opt-in (AZ_TC_SHORTCUT=1)
        import os

        self.tc_shortcut = (dev.type == "cuda" and dtype == torch.bfloat16 and net.filters == 128
                            and os.environ.get("AZ_TC_SHORTCUT", "0") == "1")
        self._side = {}

    def _side_stream(self, device):
        key = (device, torch.cuda.current_stream().cuda_stream)
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device=device)
        return self._side[key]

    def _heads_arg(self):
        from . import native

        if self._head_struct is None:
            self._head_struct = native.AzHeadWeights(
                conv_w=self.head_w32.data_ptr(), conv_b=self.head_b32.data_ptr(), policy_w=self.pfc_w_pad.data_ptr(),
                policy_b=self.pfc_b.data_ptr(), value1_w=self.v1_w_pad.data_ptr(), value1_b=self.v1_b.data_ptr(),
                value2_w=self.v2_w.data_ptr(), value2_b=self.v2_b.data_ptr())
        return self._head_struct

    @torch.no_grad()
    def forward(self, x_nhwc, priors_out=None, values_out=None, index=None, count=None, trees=None):
        """x [B, H, W, 4] -> (policy [B, A] float32 softmax, value [B] float32 tanh).  On the GPU fast path
        the results are written into priors_out / values_out when given (no extra copy kernels).
        index (int32 [B]) / count (int32 [1]), whole-net kernel only: evaluate just the rows index[:count] (the leaf list
        of az_step_gather); the other rows of priors_out / values_out keep their contents.
        trees = (engine handle, max_sims), gathered batch only: az_net_forward_trees - the engine's trees without a leaf in
        flight go on with evaluator-free simulations inside the net kernel."""
        if index is not None:
            assert self.fused_net and x_nhwc.is_cuda and priors_out is not None, "a gathered batch needs az_net_forward"
            return self._forward_fast(x_nhwc, priors_out, values_out, index, count, trees)
        if self.fast and x_nhwc.is_cuda:
            return self._forward_fast(x_nhwc, priors_out, values_out)
        B = x_nhwc.shape[0]
        x = x_nhwc.to(self.dtype).permute(0, 3, 1, 2)
        if x.is_cuda and self.dtype == torch.bfloat16:
            # other input planes / action spaces (chess: 118 planes, 1 880 actions): library stem, the same cuDNN
            # fused-epilogue tower as the fast path, heads through cuBLAS
            if self.in_pad and x_nhwc.shape[-1] == self.stem_w.shape[1]:  # not padded by the producer already
                x = F.pad(x_nhwc.to(self.dtype), (0, self.in_pad)).permute(0, 3, 1, 2)
            w_stem = self.stem_w_pad if self.in_pad else self.stem_w
            if x.shape[1] == 40 and hasattr(self, "stem_w_tail"):  # chess leaf batch without the empty history entries
                w_stem = self.stem_w_tail
            h0 = torch.cudnn_convolution_relu(x.contiguous(memory_format=torch.channels_last), w_stem, self.stem_b,
                                              (1, 1), (1, 1), (1, 1), 1)
            return self.forward_from_stem(h0.permute(0, 2, 3, 1), priors_out, values_out)
        else:
            x = F.relu_(F.conv2d(x, self.stem_w, self.stem_b, padding=1))
            for i in range(self.depth):
                w1, b1, w2, wp, b2p = self.block_params[5 * i: 5 * i + 5]
                h = F.relu_(F.conv2d(x, w1, b1, padding=1))
                y = F.conv2d(h, w2, b2p, padding=1)
                y += F.conv2d(x, wp)
                x = F.relu_(y)
            xf = x.permute(0, 2, 3, 1).float()  # [B, H, W, C]
        hd = F.relu_(F.linear(xf, self.head_w32, self.head_b32))  # 1x1 convs: [B, H, W, 3]
        return self._dense_heads_library(hd, False, priors_out, values_out)

    @torch.no_grad()
    def forward_from_stem(self, h0, priors_out=None, values_out=None):
        """The net after its stem, GPU bf16 route for shapes the fused Connect-N kernels do not cover (chess): h0
        [B, H, W, 128] bf16 NHWC (cuDNN stem or az_chess_stem) -> cuDNN fused-epilogue tower -> az_net_head_convs ->
        az_net_dense_heads (or cuBLAS)."""
        from .engine import _ptr, _stream
        from .native import AZ_DENSE_HEAD_SPLITS, check, lib

        B = h0.shape[0]
        xm = self.tower(h0)
        if self.filters == 128:  # hand-written: one pass over the tower output, float32 accumulation
            hd = torch.empty((B, self.height, self.width, 3), dtype=torch.float32, device=xm.device)
            check(lib().az_net_head_convs(_ptr(xm), _ptr(self.head_w32), _ptr(self.head_b32), B,
                                          self.height * self.width, self.filters, _ptr(hd), _stream()))
        else:
            hd = F.relu_(F.linear(xm.float(), self.head_w32, self.head_b32))
        if (self.n_actions > 128 and self.fused_dense_heads and self.filters == 128
                and self.height * self.width == 64 and self.n_actions % 4 == 0 and self.v1_w.shape[0] == 256):
            # chess: both heads' dense layers, softmax and tanh in one tcgen05 kernel
            if priors_out is None:
                priors_out = torch.empty((B, self.n_actions), dtype=torch.float32, device=hd.device)
                values_out = torch.empty(B, dtype=torch.float32, device=hd.device)
            scratch = torch.empty((B, AZ_DENSE_HEAD_SPLITS, 2), dtype=torch.float32, device=hd.device)
            check(lib().az_net_dense_heads(_ptr(hd), _ptr(self.pfc_w16), _ptr(self.pfc_b), _ptr(self.v1_w16), _ptr(self.v1_b),
                                           _ptr(self.v2_w), _ptr(self.v2_b), B, 64, self.n_actions, _ptr(priors_out),
                                           _ptr(values_out), _ptr(scratch), _stream()))
            return priors_out, values_out
        return self._dense_heads_library(hd, True, priors_out, values_out)

    def _dense_heads_library(self, hd, gpu_bf16, priors_out, values_out):
        B = hd.shape[0]
        p = hd[..., :2].reshape(B, -1)
        v = hd[..., 2].reshape(B, -1)
        if gpu_bf16 and self.n_actions > 128:
            # wide policy layer without the fused kernel: bf16 operands on the tensor cores, float32 accumulation and softmax
            logits = F.linear(p.to(torch.bfloat16), self.pfc_w16[: self.n_actions]).float() + self.pfc_b
        else:
            logits = F.linear(p, self.pfc_w, self.pfc_b)
        policy = torch.softmax(logits, dim=-1)
        value = torch.tanh(F.linear(F.relu_(F.linear(v, self.v1_w, self.v1_b)), self.v2_w, self.v2_b)).reshape(B)
        if priors_out is not None:
            priors_out.copy_(policy)
            values_out.copy_(value)
            return priors_out, values_out
        return policy, value

    def _net_heads_arg(self):
        from . import native

        if self._net_heads is None:
            self._net_heads = native.AzNetHeadParams(
                conv_w=self.head_w32.data_ptr(), conv_b=self.head_b32.data_ptr(), policy_w=self.pfc_w.data_ptr(),
                policy_b=self.pfc_b.data_ptr(), value1_w=self.v1_w.data_ptr(), value1_b=self.v1_b.data_ptr(),
                value2_w=self.v2_w.data_ptr(), value2_b=self.v2_b.data_ptr())
        return self._net_heads

    def _forward_fast(self, x_nhwc, priors_out, values_out, index=None, count=None, trees=None):
        import ctypes

        from .engine import _ptr, _stream
        from .native import check, lib

        B, H, W = x_nhwc.shape[0], self.height, self.width
        x_nhwc = x_nhwc.to(torch.bfloat16).contiguous()
        if self.fused_net:
            # the whole forward pass in one kernel: planes in, priors and values out
            if priors_out is None:
                priors_out = torch.empty((B, self.n_actions), dtype=torch.float32, device=x_nhwc.device)
                values_out = torch.empty(B, dtype=torch.float32, device=x_nhwc.device)
            if index is not None and trees is not None:
                check(lib().az_net_forward_trees(_ptr(x_nhwc), _ptr(self.net_img), _ptr(self.stem_b32), _ptr(self.tower_bias),
                                                 ctypes.byref(self._net_heads_arg()), _ptr(index), _ptr(count), B, H, W,
                                                 self.filters, self.depth, self.n_actions, self.tower_layout,
                                                 _ptr(priors_out), _ptr(values_out), trees[0], int(trees[1]), _stream()))
                return priors_out, values_out
            if index is not None:
                check(lib().az_net_forward_gathered(_ptr(x_nhwc), _ptr(self.net_img), _ptr(self.stem_b32), _ptr(self.tower_bias),
                                                    ctypes.byref(self._net_heads_arg()), _ptr(index), _ptr(count), B, H, W,
                                                    self.filters, self.depth, self.n_actions, self.tower_layout,
                                                    _ptr(priors_out), _ptr(values_out), _stream()))
                return priors_out, values_out
            check(lib().az_net_forward(_ptr(x_nhwc), _ptr(self.net_img), _ptr(self.stem_b32), _ptr(self.tower_bias),
                                       ctypes.byref(self._net_heads_arg()), B, H, W, self.filters, self.depth, self.n_actions,
                                       self.tower_layout, _ptr(priors_out), _ptr(values_out), _stream()))
            return priors_out, values_out
        h0 = torch.empty((B, H, W, self.filters), dtype=torch.bfloat16, device=x_nhwc.device)
        if self.tc_stem:
            check(lib().az_net_stem_tc(_ptr(x_nhwc), _ptr(self.stem_w16_k), _ptr(self.stem_b32), B, H, W, self.filters,
                                       _ptr(h0), _stream()))
        else:
            check(lib().az_net_stem(_ptr(x_nhwc), _ptr(self.stem_w32), _ptr(self.stem_b32), B, H, W, self.filters,
                                    _ptr(h0), _stream()))
        xm = self.tower(h0)
        if priors_out is None:
            priors_out = torch.empty((B, self.n_actions), dtype=torch.float32, device=xm.device)
            values_out = torch.empty(B, dtype=torch.float32, device=xm.device)
        hw = self._heads_arg()
        if self.split_heads:  # 128-bit-load head convolutions, then the dense layers
            hd = torch.empty((B, H * W, 3), dtype=torch.float32, device=xm.device)
            check(lib().az_net_head_convs(_ptr(xm), _ptr(self.head_w32), _ptr(self.head_b32), B, H * W, self.filters, _ptr(hd),
                                          _stream()))
            check(lib().az_net_heads_dense(_ptr(hd), ctypes.byref(hw), B, H * W, self.n_actions, _ptr(priors_out),
                                           _ptr(values_out), _stream()))
        else:
            check(lib().az_net_heads(_ptr(xm), ctypes.byref(hw), B, H * W, self.filters, self.n_actions, _ptr(priors_out),
                                     _ptr(values_out), _stream()))
        return priors_out, values_out

    @torch.no_grad()
    def tower(self, h0):
        """The residual tower on the stem output h0 [B, H, W, 128] bf16 (NHWC) -> [B, H, W, 128] bf16, NHWC-contiguous:
        az_net_tower (one hand-written tcgen05 kernel, activations resident on chip) or, where that does not apply,
        4 x (cuDNN conv+bias+ReLU, cuDNN 1x1 shortcut, cuDNN conv+shortcut+bias+ReLU)."""
        if self.fused_tower and h0.is_cuda:
            from .engine import _ptr, _stream
            from .native import check, lib

            h0 = h0 if h0.is_contiguous() else h0.contiguous()
            out = torch.empty_like(h0)
            check(lib().az_net_tower(_ptr(h0), _ptr(self.tower_img), _ptr(self.tower_bias), h0.shape[0], self.height,
                                     self.width, self.filters, self.depth, self.tower_layout, _ptr(out), _stream()))
            return out
        return self.tower_library(h0)

    @torch.no_grad()
    def tower_library(self, h0):
        """The same tower through cuDNN (12 launches): the comparison arm of az_net_tower and the route for boards
        its tiling does not fit (9x9)."""
        x = h0.permute(0, 3, 1, 2)  # logical NCHW over NHWC memory (channels_last)
        one = (1, 1)
        cur = torch.cuda.current_stream()
        side = self._side_stream(x.device) if self.overlap_shortcut else None
        for i in range(self.depth):
            w1, b1, w2, wp, b2p = self.block_params[5 * i: 5 * i + 5]
            if side is not None:
                # the bandwidth-bound 1x1 shortcut runs beside the compute-bound 3x3 on a forked stream
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    p = F.conv2d(x, wp)
                p.record_stream(cur)
                h = torch.cudnn_convolution_relu(x, w1, b1, one, one, one, 1)
                cur.wait_stream(side)
            elif self.tc_shortcut:
                h = torch.cudnn_convolution_relu(x, w1, b1, one, one, one, 1)       # conv + bias + ReLU
                p = self._shortcut_tc(x, wp)                                         # projection shortcut (tcgen05)
            else:
                h = torch.cudnn_convolution_relu(x, w1, b1, one, one, one, 1)       # conv + bias + ReLU
                p = F.conv2d(x, wp)                                                  # projection shortcut
            x = torch.cudnn_convolution_add_relu(h, w2, p, 1.0, b2p, one, one, one, 1)  # conv + shortcut + bias + ReLU
        xm = x.permute(0, 2, 3, 1)
        return xm if xm.is_contiguous() else xm.contiguous()

    def _shortcut_tc(self, x, wp):
        """x: logical NCHW over NHWC memory, bf16; wp [128, 128, 1, 1] -> the same kind of tensor."""
        from .engine import _ptr, _stream
        from .native import check, lib

        xn = x.permute(0, 2, 3, 1)
        if not xn.is_contiguous():
            xn = xn.contiguous()
        out = torch.empty_like(xn)
        check(lib().az_net_conv1x1(_ptr(xn), _ptr(wp), xn.numel() // self.filters, self.filters, _ptr(out), _stream()))
        return out.permute(0, 3, 1, 2)

    @torch.no_grad()
    def chess_stem(self, positions, tc=True):
        """positions int64 [B, 8] (az_chess_pos) on the GPU -> stem output [B, 8, 8, 128] bf16: az_chess_stem_tc
        (tcgen05) or az_chess_stem (mma.sync)."""
        from .engine import _ptr, _stream
        from .native import check, lib

        B = positions.shape[0]
        out = torch.empty((B, 8, 8, self.filters), dtype=torch.bfloat16, device=positions.device)
        if tc:
            check(lib().az_chess_stem_tc(_ptr(positions), B, _ptr(self.chess_stem_w16), _ptr(self.chess_stem_map), _ptr(out),
                                         _stream()))
        else:
            check(lib().az_chess_stem(_ptr(positions), B, _ptr(self.chess_stem_w), _ptr(self.chess_stem_map), _ptr(out), _stream()))
        return out

    def load_from(self, net: PolicyValueNet):
        """Refreshes the folded weights in place (after a training step / weight broadcast): the CUDA
        graph that captured forward() keeps replaying with the new values."""
        fresh = InferenceNet(net, dtype=self.dtype, device=self.stem_w.device)
        with torch.no_grad():
            for dst, src in zip(self.parameters(), fresh.parameters()):
                dst.copy_(src)

    def flat_weights(self):
        return torch.cat([p.detach().reshape(-1).float() for p in self.parameters()])


def randomise_bn(net: PolicyValueNet, seed=0):
    """Gives the BN layers non-trivial statistics so that folding is actually exercised in tests."""
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5, generator=g)
            m.bias.data.uniform_(-0.2, 0.2, generator=g)
            m.running_mean.data.uniform_(-0.2, 0.2, generator=g)
            m.running_var.data.uniform_(0.5, 1.5, generator=g)
    return net


def glorot_limit(fan_in, fan_out):
    return math.sqrt(6.0 / (fan_in + fan_out))
