"""CPU emulation of csrc/tc_conv2.cu from the launch geometry the library itself computes (enerf_tc_conv2_plan, no GPU
needed): tile / halo / TMA box coordinates (incl. element-stride-2 phase tiles and zero fill), the SWIZZLE_{32,64,128}B
shared-memory image at BYTE-ADDRESS level (both the box write and the operand read XOR address bits [4,4+B) with bits
[7,7+B)), tap start-address offsets, K-steps (+32 B), K-blocks, the weight pack layout (packing.pack_tc_conv /
pack_tc_deconv), the accumulator-row -> voxel mapping, the kx-fold shift and the transposed convolution's pixel shuffle.
The emulated layer must equal torch's convolution.  What this cannot cover is the hardware contract itself (that the
tensor core and the TMA unit swizzle on absolute address bits): tests/test_parity_gpu.py::test_tcgen05_swizzled_tma_operand_
with_row_offsets pins that on the B200.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from enerf_b200 import capi, packing

TC_PLAIN, TC_HEAD, TC_DECONV, TC_SINGLE = 0, 1, 2, 3


def _swz(addr, row_bytes):
    bits = {128: 7, 64: 3, 32: 1}[row_bytes]
    return addr ^ (((addr >> 7) & bits) << 4)


def emulate(plan, kind, KD, KH, stride, cin, cout, mode, relu, x, wpack, bias, skip):
    """x: (Di,Hi,Wi,cin) float32 numpy (the layer's input), returns the layer output(s) as the kernel would write them."""
    P = plan
    Dn, Hn, Wn = (x.shape[0] // P["sz"], x.shape[1] // P["sy"], x.shape[2] // P["sx"]) if kind == 0 else x.shape[:3]
    kbc, N, n_mt, n_taps = P["kbc"], P["N"], P["n_mt"], P["n_taps"]
    rb = kbc * 4
    ksteps = kbc // 8
    npix = P["IZ"] * P["IY"] * P["IX"]
    assert P["box_bytes"] == npix * rb and P["slot_bytes"] % 1024 == 0 and P["phase_bytes"] % 1024 == 0
    wp = wpack.reshape(cin // 8, n_taps, 2, N, 4)           # [K-stage][tap][chunk][n][4]
    if kind == 1:
        out = skip.copy()
    elif mode == TC_HEAD:
        out, out2 = np.full((Dn, Hn, Wn, 8), np.nan, np.float32), np.full((Dn, Hn, Wn), np.nan, np.float32)
    elif mode == TC_SINGLE:
        out = np.full((Dn, Hn, Wn), np.nan, np.float32)
    else:
        out = np.full((Dn, Hn, Wn, cout), np.nan, np.float32)
    Di, Hi, Wi = x.shape[:3]
    base = 3 * 1024                                          # any 1024-aligned slot base
    for tile in range(P["n_tiles"]):
        bx, by, bz = tile % P["nx"], (tile // P["nx"]) % P["ny"], tile // (P["nx"] * P["ny"])
        acc = np.zeros((n_mt * 128 + 4, N), np.float64)
        for kb in range(P["n_kb"]):
            smem = np.full((base + P["slot_bytes"]) // 4 + 64, np.nan, np.float32)     # NaN = never written (garbage rows)
            # ---- producer: one TMA box per phase, element stride = conv stride, out-of-bounds = 0 ----
            x0, y0, z0 = bx * P["TX"] - P["ox"], by * P["TY"] - P["oy"], bz * P["TZ"] - P["oz"]
            for ph in range(P["n_phases"]):
                rx, ry, rz = ph % P["sx"], (ph // P["sx"]) % P["sy"], ph // (P["sx"] * P["sy"])
                cz0, cy0, cx0 = P["sz"] * z0 + rz, P["sy"] * y0 + ry, P["sx"] * x0 + rx
                for z in range(P["IZ"]):
                    for y in range(P["IY"]):
                        for xx in range(P["IX"]):
                            gz, gy, gx = cz0 + z * P["sz"], cy0 + y * P["sy"], cx0 + xx * P["sx"]
                            inb = 0 <= gz < Di and 0 <= gy < Hi and 0 <= gx < Wi
                            v = x[gz, gy, gx, kb * kbc:(kb + 1) * kbc] if inb else np.zeros(kbc, np.float32)
                            row = (z * P["IY"] + y) * P["IX"] + xx
                            for c16 in range(rb // 16):
                                a = _swz(base + ph * P["phase_bytes"] + row * rb + c16 * 16, rb)
                                smem[a // 4:a // 4 + 4] = v[c16 * 4:c16 * 4 + 4]
            # ---- MMA: M-tiles x K-steps x taps, operand start = slot + (128 m) rows + tap offset + 32 B per K-step ----
            for m in range(n_mt):
                for ks in range(ksteps):
                    st = kb * ksteps + ks
                    for tp in range(n_taps):
                        start = base + (m * 128 * rb) + P["tap_off"][tp] * 16 + ks * 32
                        A = np.empty((128, 8), np.float64)
                        for r in range(128):
                            for j in range(2):
                                a = _swz(start + r * rb + j * 16, rb)
                                A[r, 4 * j:4 * j + 4] = smem[a // 4:a // 4 + 4]
                        B = wp[st, tp].transpose(1, 0, 2).reshape(N, 8).astype(np.float64)     # [n][k = chunk*4 + i]
                        contrib = A @ B.T
                        acc[m * 128:(m + 1) * 128] += np.where(np.isnan(contrib), np.nan, contrib)
        # ---- epilogue: row q -> halo position -> voxel ----
        plane = P["IY"] * P["IX"]
        for q in range(n_mt * 128):
            z, rem = divmod(q, plane)
            y, xx = divmod(rem, P["IX"])
            gz, gy, gx = bz * P["TZ"] + z, by * P["TY"] + y, bx * P["TX"] + xx
            if not (z < P["TZ"] and y < P["TY"] and xx < P["TX"] and gz < Dn and gy < Hn and gx < Wn):
                continue
            if P["fold"]:
                C = {TC_PLAIN: cout, TC_HEAD: 9, TC_SINGLE: 1}[mode]
                v = acc[q, 0:C] + acc[q + 1, C:2 * C] + acc[q + 2, 2 * C:3 * C]
            else:
                v = acc[q]
            assert not np.isnan(v[:1]).any(), "a valid output read a row that was never loaded"
            if kind == 1:
                for e in range(8):
                    o = (2 * gz + (e >> 2), 2 * gy + ((e >> 1) & 1), 2 * gx + (e & 1))
                    out[o] = skip[o] + (v[e * cout:(e + 1) * cout] + bias)
            elif mode == TC_HEAD:
                out[gz, gy, gx], out2[gz, gy, gx] = v[:8], v[8]
            elif mode == TC_SINGLE:
                out[gz, gy, gx] = v[0]
            else:
                r = v[:cout] + bias
                out[gz, gy, gx] = np.maximum(r, 0) if relu else r
    return (out, out2) if (kind == 0 and mode == TC_HEAD) else out


CASES = [
    # kind KD KH stride cin cout mode relu (row grid D,H,W)  n_sm
    (0, 1, 3, 1, 32, 8, TC_PLAIN, 0, (2, 9, 40), 2),        # smooth0-like: one 128-byte box per pixel row (SWIZZLE_128B)
    (0, 1, 3, 1, 16, 16, TC_PLAIN, 1, (1, 10, 36), 148),    # 64-byte rows, shrunk tile (small layer on a big device)
    (0, 1, 3, 1, 8, 8, TC_PLAIN, 1, (1, 17, 33), 2),        # 32-byte rows, ragged extents
    (0, 1, 1, 1, 32, 32, TC_PLAIN, 0, (2, 9, 33), 2),       # 1x1
    (0, 3, 3, 1, 16, 8, TC_PLAIN, 1, (5, 6, 34), 2),        # CostRegNet conv0: kx folded into N
    (0, 3, 3, 1, 32, 8, TC_PLAIN, 1, (4, 5, 33), 2),        # MinCostRegNet conv0: folded, 32 channels
    (0, 3, 3, 1, 16, 16, TC_PLAIN, 1, (3, 9, 20), 2),       # conv2: 27 taps unfolded
    (0, 3, 3, 1, 8, 9, TC_HEAD, 0, (3, 5, 35), 2),          # head (feat + prob), folded
    (0, 3, 3, 1, 8, 1, TC_SINGLE, 0, (3, 5, 35), 2),        # depth head only, folded
    (1, 3, 3, 1, 16, 8, TC_DECONV, 0, (2, 4, 20), 2),       # conv11: sub-pixel transposed conv
    (0, 3, 3, 2, 8, 16, TC_PLAIN, 1, (2, 4, 18), 2),        # conv1: stride 2 = 8 phase boxes with element stride 2
    (0, 1, 5, 2, 8, 16, TC_PLAIN, 1, (1, 8, 20), 2),        # FeatureNet conv1.0: 5x5 stride 2 = 4 phase boxes
]


@pytest.mark.parametrize("case", CASES)
def test_tc_conv2_geometry_emulated_on_cpu(case):
    kind, KD, KH, stride, cin, cout, mode, relu, (D, H, W), n_sm = case
    g = torch.Generator().manual_seed(cin * 31 + cout + KD)
    fold = packing.tc_fold_kx(KD, KH, stride, cout, single=(mode == TC_SINGLE), head=(mode == TC_HEAD), cin=cin) if kind == 0 else False
    plan = capi.tc_conv2_plan(kind, KD, KH, stride, cin, cout, mode, D, H, W, fold, n_sm=n_sm)
    assert plan["fold"] == int(fold)
    sz = stride if (KD > 1 and kind == 0) else 1
    xin = torch.randn(1, cin, D * sz, H * (stride if kind == 0 else 1), W * (stride if kind == 0 else 1), generator=g)
    if kind == 0:
        w = torch.randn(cout, cin, KD, KH, KH, generator=g) / (cin * KD * KH * KH) ** 0.5
        b = (torch.randn(cout, generator=g) * 0.1) if mode == TC_PLAIN else None
        ref = F.conv3d(xin, w, b, (sz, stride, stride), (KD // 2, KH // 2, KH // 2))
        if relu:
            ref = F.relu(ref)
        wp = packing.pack_tc_conv(packing._taps_cin_cout(w), fold_kx=fold)
        skip = None
    else:
        w = torch.randn(cin, cout, 3, 3, 3, generator=g) / (cin * 27 / 8) ** 0.5
        b = torch.randn(cout, generator=g) * 0.1
        skip_t = torch.randn(1, cout, 2 * D, 2 * H, 2 * W, generator=g)
        ref = skip_t + F.conv_transpose3d(xin, w, b, stride=2, padding=1, output_padding=1)
        wp = packing.pack_tc_deconv(w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout))
        skip = skip_t[0].permute(1, 2, 3, 0).contiguous().numpy()
    assert wp.numel() * 4 == plan["w_bytes"]
    # the packed weights are TF32-rounded; round the reference's too by unpacking is overkill: compare at 2e-3 like the GPU test
    got = emulate(plan, kind, KD, KH, stride, cin, cout, mode, relu, xin[0].permute(1, 2, 3, 0).contiguous().numpy(), wp.numpy(),
                  b.numpy() if b is not None else None, skip)
    refc = ref[0].permute(1, 2, 3, 0).numpy()
    if kind == 0 and mode == TC_HEAD:
        got = np.concatenate([got[0], got[1][..., None]], -1)
    elif mode == TC_SINGLE:
        got = got[..., None]
    assert not np.isnan(got).any(), "some output voxel was never written"
    err = np.abs(got - refc).max()
    assert err < 2e-3 * max(1.0, np.abs(refc).max()), f"max abs err {err}"


@pytest.mark.parametrize("NB", [2, 3, 4])
@pytest.mark.parametrize("n_mt", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_fold_epilogue_exchange_schedule(n_mt, NB):
    """The kx-fold epilogue of csrc/tc_conv2.cu reads M-tiles in batches of NB (+ the first M-tile of the next batch as a look-ahead,
    only for its rows 0 / 1), publishes rows 0 / 1 of every 32-row group (entry m * 4 + g) to the exchange buffer, passes ONE named
    barrier per batch and then forms the outputs of the batch, which read entry m * 4 + g + 1.  Model of that schedule: every entry an
    output needs was published at or before its batch's barrier (or belongs to rows past the tile, which no valid output uses), and
    no entry is ever published twice (so no thread writes an entry another may still be reading)."""
    published = {}                      # entry -> batch index that published it
    for bi, m0 in enumerate(range(0, n_mt, NB)):
        nb = min(NB, n_mt - m0)
        look = m0 + NB < n_mt
        for b in range(nb):
            if b > 0 or m0 == 0:        # a batch's first M-tile was published as the previous batch's look-ahead
                for g in range(4):
                    e = (m0 + b) * 4 + g
                    assert e not in published, f"entry {e} published twice"
                    published[e] = bi
        if look:
            for g in range(4):
                e = (m0 + NB) * 4 + g
                assert e not in published, f"look-ahead entry {e} published twice"
                published[e] = bi
        # ---- barrier of batch bi ----
        for b in range(nb):
            m = m0 + b
            for g in range(4):
                need = m * 4 + g + 1    # rows 0, 1 of the next 32-row group
                if need == n_mt * 4:    # the group after the tile's last one: garbage by construction, rows past the accumulator
                    continue            # (n_mt is sized with the fold's 2 extra rows, so no valid output reads it)
                assert need in published and published[need] <= bi, f"M-tile {m} group {g} reads entry {need} before it is published"
    assert sorted(published) == list(range(n_mt * 4)), "every group's rows 0 / 1 are published exactly once"
