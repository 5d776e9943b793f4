"""CPU check of the fused tower's data layout (csrc/az_tower.cu, az_b200/net.py pack_tower_weights): a numpy model of
exactly what the kernel's descriptors address - one linear row space with zero padding, rows ordered (y, position, x),
left / right masked copies for the dx = -1 / +1 taps, weight stages decoded by their stream position - must equal the
reference's residual tower (model/tensorflow/base_layers.py:85-125) computed by torch convolutions.  No GPU needed:
this pins the index arithmetic; tests/test_gpu_tower.py pins the kernel itself."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from az_b200.net import pack_tower_weights

PAD, TILE, NBUF = 22, 128, 4
BUF = TILE + PAD
ROWS = PAD + NBUF * BUF


def emulate_tower(x, img, bias, H, W, depth):
    """x float32 [n, H, W, 128]; img float32 flat stage stream; bias [depth, 2, 128] -> [n, H, W, 128]."""
    n, cells = x.shape[0], H * W
    ppt = TILE // cells
    rowstride = ppt * W
    assert rowstride + 1 <= PAD
    stages = img.reshape(depth, 38, 8, 128, 8)  # block, stage, chunk, cout, e

    def b_operand(b, s):  # -> [cout][64] of stage s of block b
        return stages[b, s].transpose(1, 0, 2).reshape(128, 64)

    out = np.zeros_like(x)
    row0 = [PAD + i * BUF for i in range(NBUF)]
    for tile in range((n + ppt - 1) // ppt):
        space = np.zeros((ROWS, 128), np.float32)

        def store(centre, vals):  # vals [TILE][128] for the live rows
            for r in range(ppt * cells):
                xcol = (r % rowstride) % W
                space[row0[centre] + r] = vals[r]
                space[row0[1] + r] = 0 if xcol == W - 1 else vals[r]
                space[row0[2] + r] = 0 if xcol == 0 else vals[r]

        tile_in = np.zeros((TILE, 128), np.float32)
        for r in range(ppt * cells):
            y, rem = divmod(r, rowstride)
            p, xc = divmod(rem, W)
            if tile * ppt + p < n:
                tile_in[r] = x[tile * ppt + p, y, xc]
        store(0, tile_in)
        for b in range(depth):
            def conv3(centre, s0):
                acc = np.zeros((TILE, 128), np.float32)
                for tap in range(9):
                    dy, dx = tap // 3 - 1, tap % 3 - 1
                    buf = 1 if dx < 0 else (2 if dx > 0 else centre)
                    start = row0[buf] + dy * rowstride + dx
                    a = space[start:start + TILE]
                    for kb in range(2):
                        acc += a[:, kb * 64:(kb + 1) * 64] @ b_operand(b, s0 + kb * 9 + tap).T
                return acc
            h = np.maximum(conv3(0, 0) + bias[b, 0], 0)
            a0 = space[row0[0]:row0[0] + TILE]
            sc = a0[:, :64] @ b_operand(b, 18).T + a0[:, 64:] @ b_operand(b, 19).T
            store(3, bf16_round(h))
            yv = np.maximum(conv3(3, 20) + sc + bias[b, 1], 0)
            store(0, bf16_round(yv))
        for r in range(ppt * cells):
            y, rem = divmod(r, rowstride)
            p, xc = divmod(rem, W)
            if tile * ppt + p < n:
                out[tile * ppt + p, y, xc] = space[row0[0] + r]
    return out


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


def torch_tower(x, blocks):
    t = torch.from_numpy(x).permute(0, 3, 1, 2)
    for w1, b1, w2, wp, b2p in blocks:
        h = F.relu(F.conv2d(t, w1, b1, padding=1)).to(torch.bfloat16).float()
        t = F.relu(F.conv2d(h, w2, b2p, padding=1) + F.conv2d(t, wp)).to(torch.bfloat16).float()
    return t.permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("H,W,n,depth", [(6, 7, 7, 2), (8, 8, 3, 1), (7, 7, 5, 1), (9, 9, 2, 2)])
def test_layout_model_equals_the_convolutions(H, W, n, depth):
    g = torch.Generator().manual_seed(H * 100 + W)
    blocks = []
    for _ in range(depth):
        w1 = (torch.randn(128, 128, 3, 3, generator=g) * 0.03).to(torch.bfloat16).float()
        w2 = (torch.randn(128, 128, 3, 3, generator=g) * 0.03).to(torch.bfloat16).float()
        wp = (torch.randn(128, 128, 1, 1, generator=g) * 0.08).to(torch.bfloat16).float()
        blocks.append((w1, torch.randn(128, generator=g) * 0.1, w2, wp, torch.randn(128, generator=g) * 0.1))
    x = bf16_round(torch.rand(n, H, W, 128, generator=g).numpy())
    img, bias = pack_tower_weights(blocks)
    assert img.numel() == depth * 38 * 8192 and bias.shape == (depth, 2, 128)
    got = emulate_tower(x, img.float().numpy(), bias.numpy(), H, W, depth)
    want = torch_tower(x, blocks)
    # identical operands, float32 accumulation in a different order, bf16 rounding of the activations at the same places
    assert np.abs(got - want).max() <= 2 ** -6 * max(1.0, np.abs(want).max())
    assert np.mean(np.abs(got - want) > 1e-4) < 0.02


def test_stem_stage_layout():
    """pack_stem_weights: tap t of the 4-plane stem is one K = 16 slice (planes in K 0-3) of stage t // 4 at chunk columns
    2 * (t % 4), 2 * (t % 4) + 1; read back through the same window arithmetic it must give the stem convolution."""
    from az_b200.net import pack_stem_weights

    g = torch.Generator().manual_seed(3)
    H, W, n = 6, 7, 5
    w = (torch.randn(128, 4, 3, 3, generator=g) * 0.3).to(torch.bfloat16).float()
    b = torch.randn(128, generator=g) * 0.1
    x = torch.randint(0, 2, (n, H, W, 4), generator=g).float()
    img = pack_stem_weights(w).float().numpy().reshape(3, 8, 128, 8)  # stage, chunk column, cout, e
    cells, ppt = H * W, TILE // (H * W)
    rowstride = ppt * W
    want = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1)).permute(0, 2, 3, 1).numpy()
    for tile in range((n + ppt - 1) // ppt):
        space = np.zeros((ROWS, 16), np.float32)  # chunk columns 0-1 of the row space: K = 16
        row0 = [PAD + i * BUF for i in range(3)]
        for r in range(ppt * cells):
            y, rem = divmod(r, rowstride)
            p, xc = divmod(rem, W)
            v = x[tile * ppt + p, y, xc].numpy() if tile * ppt + p < n else np.zeros(4, np.float32)
            space[row0[0] + r, :4] = v
            space[row0[1] + r, :4] = 0 if xc == W - 1 else v
            space[row0[2] + r, :4] = 0 if xc == 0 else v
        acc = np.zeros((TILE, 128), np.float32)
        for tap in range(9):
            dy, dx = tap // 3 - 1, tap % 3 - 1
            start = row0[1 if dx < 0 else (2 if dx > 0 else 0)] + dy * rowstride + dx
            bmat = np.concatenate([img[tap // 4, 2 * (tap % 4)], img[tap // 4, 2 * (tap % 4) + 1]], axis=1)  # [cout][16]
            acc += space[start:start + TILE] @ bmat.T
        out = np.maximum(acc + b.numpy(), 0)
        for r in range(ppt * cells):
            y, rem = divmod(r, rowstride)
            p, xc = divmod(rem, W)
            if tile * ppt + p < n:
                assert np.allclose(out[r], want[tile * ppt + p, y, xc], atol=1e-5)


def test_pair_layout_is_the_single_layout_split_by_output_channel():
    """pack_*_weights(pair=True): stage s, half h, chunk c, row n, element e == the single-CTA image at stage s, chunk c,
    output channel 64 h + n, element e - CTA h of a pair streams bytes [s * 16 KB + h * 8 KB, + 8 KB)."""
    from az_b200.net import pack_stem_weights

    g = torch.Generator().manual_seed(5)
    blocks = [((torch.randn(128, 128, 3, 3, generator=g)), torch.zeros(128), torch.randn(128, 128, 3, 3, generator=g),
               torch.randn(128, 128, 1, 1, generator=g), torch.zeros(128))]
    one, _ = pack_tower_weights(blocks)
    two, _ = pack_tower_weights(blocks, pair=True)
    a = one.float().reshape(38, 8, 128, 8)
    b = two.float().reshape(38, 2, 8, 64, 8)
    for h in range(2):
        assert torch.equal(b[:, h], a[:, :, 64 * h: 64 * h + 64])
    w = torch.randn(128, 4, 3, 3, generator=g)
    s1 = pack_stem_weights(w).float().reshape(3, 8, 128, 8)
    s2 = pack_stem_weights(w, pair=True).float().reshape(3, 2, 8, 64, 8)
    for h in range(2):
        assert torch.equal(s2[:, h], s1[:, :, 64 * h: 64 * h + 64])
