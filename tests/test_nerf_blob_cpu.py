"""CPU check of the factorised per-sample MLP exactly as csrc/render_rays_ws.cu slices the packed weight blob
(packing.pack_nerf_tc): every B operand is decoded from the blob at the float offsets / chunk ranges the kernel's MMA
descriptors use, the GEMMs are grouped as the kernel groups them (shared part once per point + per-view part, the bias
riding in the constant-1 input column), and the result must equal the oracle's NeRF.forward (nerf.py:29-43,74-89).
This pins: the blob layout (struct W), the K-range split of lr0 / color.0 / global_fc, and the bias-in-B trick."""
import pytest
import torch
import torch.nn.functional as F

from enerf_b200 import config as bcfg, packing, synthetic
from oracle import enerf_oracle as O

# struct W of csrc/render_rays_ws.cu (float offsets)
OFF = {}
o = 0
for name, n in (("bg_view", 4 * 32 * 4), ("bg_shared", 6 * 32 * 4), ("bfc", 8 * 16 * 4), ("b0", 6 * 64 * 4), ("bc_shared", 22 * 64 * 4), ("bc_view", 4 * 64 * 4),
                ("v_view_w", 48), ("v_view_b", 12), ("v_bg", 32), ("v_wa", 32), ("v_ba", 4), ("v_bf", 16), ("v_b0", 64), ("v_ws", 64), ("v_bs", 4),
                ("v_bc", 64), ("v_w2", 64), ("v_b2", 4)):
    OFF[name] = o
    o += n
TOTAL = o


def B(blob, name, chunk0, n_chunks, N):
    """The (4*n_chunks, N) matrix the tensor core sees for chunks [chunk0, chunk0+n_chunks) of a K-major [K/4][N][4] operand."""
    base = OFF[name] + chunk0 * N * 4
    return blob[base:base + n_chunks * N * 4].view(n_chunks, N, 4).permute(0, 2, 1).reshape(n_chunks * 4, N)


@pytest.mark.parametrize("S,viewdir", [(3, True), (2, True), (3, False)])
def test_factorised_mlp_from_blob_matches_oracle(S, viewdir):
    assert TOTAL == 10392
    cfg = bcfg.set_cfg(bcfg.make_cfg(volume_planes=[8, 8], render_if=[False, True], viewdir_agg=viewdir))
    from enerf_b200.network import Network
    torch.manual_seed(0)
    net = Network().eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    blob = packing.pack_nerf_tc(sd, "nerf_1", 11, viewdir, "cpu")
    g = torch.Generator().manual_seed(S)
    n = 257
    vox = torch.randn(1, n, 8, generator=g)
    f = torch.randn(1, n, S, 15, generator=g)
    ref = O.nerf(sd, "nerf_1", vox, f, viewdir)[0]                                   # (n, 4)

    r = lambda t: packing.tf32_round(t)                                             # operands are TF32 (weights pre-rounded in the blob)
    fs, dirs = f[0, :, :, :11], f[0, :, :, 11:]
    vw = blob[OFF["v_view_w"]:OFF["v_view_w"] + 48].view(4, 12)[:, :11]
    vb = blob[OFF["v_view_b"]:OFF["v_view_b"] + 11]
    gs = fs + (F.relu(dirs @ vw + vb) if viewdir else 0.0)
    var, mean = gs.var(dim=1, unbiased=True), gs.mean(dim=1)
    ones, z1 = torch.ones(n, 1), torch.zeros(n, 1)
    # batch 1: global_fc shared [var(11),0 | mean(11),0] (K=24) + per view [g(11),1,0,0,0,0] (K=16)
    g1s = r(torch.cat([var, z1, mean, z1], 1)) @ B(blob, "bg_shared", 0, 6, 32)
    im_num, logits, hs = 0, [], []
    for s in range(S):
        a = r(torch.cat([gs[:, s], ones, torch.zeros(n, 4)], 1))
        h = F.relu(g1s + a @ B(blob, "bg_view", 0, 4, 32))
        hs.append(h)
        logits.append(F.relu(h @ blob[OFF["v_wa"]:OFF["v_wa"] + 32] + blob[OFF["v_ba"]]))
    w = torch.softmax(torch.stack(logits, 1), dim=1)
    im = sum(w[:, s:s + 1] * hs[s] for s in range(S))
    img = F.relu(r(im) @ B(blob, "bfc", 0, 8, 16) + blob[OFF["v_bf"]:OFF["v_bf"] + 16])
    # lr0 = vox columns (chunks 0,1 of b0) + img columns (chunks 2..5)
    x = F.relu(r(vox[0]) @ B(blob, "b0", 0, 2, 64) + r(img) @ B(blob, "b0", 2, 4, 64) + blob[OFF["v_b0"]:OFF["v_b0"] + 64])
    sigma = F.softplus(x @ blob[OFF["v_ws"]:OFF["v_ws"] + 64] + blob[OFF["v_bs"]])
    # color.0 shared = x (chunks 0..15) + vox (16,17) + img (18..21) of bc_shared; per view [f(11), dir(4), 1] x bc_view
    cs = r(x) @ B(blob, "bc_shared", 0, 16, 64) + r(vox[0]) @ B(blob, "bc_shared", 16, 2, 64) + r(img) @ B(blob, "bc_shared", 18, 4, 64)
    cl = []
    for s in range(S):
        a = r(torch.cat([fs[:, s], dirs[:, s], ones], 1))
        hc = F.relu(cs + a @ B(blob, "bc_view", 0, 4, 64))
        cl.append(F.relu(hc @ blob[OFF["v_w2"]:OFF["v_w2"] + 64] + blob[OFF["v_b2"]]))
    wc = torch.softmax(torch.stack(cl, 1), dim=1)
    rgb = sum(wc[:, s:s + 1] * fs[:, s, 8:11] for s in range(S))
    got = torch.cat([rgb, sigma[:, None]], 1)
    err = (got - ref).abs().max().item()
    assert err < 5e-3 * max(1.0, ref.abs().max().item()), err      # TF32 operand rounding on unit-variance random features
