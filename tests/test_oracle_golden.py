"""Pins the CPU oracle (oracle/enerf_oracle.py) against outputs of the unmodified reference
(tests/golden/*.pt, minted by oracle/make_golden.py from /root/reference @5a084e9).
The reference's own tests hold no vectors for this path."""
import torch

from oracle import enerf_oracle as O

# fp32 noise floor of the reference against itself is ~5e-6 on rgb; the
# oracle uses the same torch ops in the same order, so it should sit at or below that.
TOL = dict(rgb=2e-5, depth=2e-5, weights=2e-5, depth_mvs=2e-5, std=2e-5)


def test_oracle_matches_reference_outputs(golden):
    torch.set_num_threads(8)
    with torch.no_grad():
        out, mid = O.forward(golden["state_dict"], golden["cfg"], golden["batch"], intermediates=True, human=golden.get("human", False))
    assert set(out) == set(golden["out"])
    for k, ref in golden["out"].items():
        err = (out[k] - ref).abs().max().item()
        tol = TOL[k.split("_level")[0]]
        assert out[k].shape == ref.shape
        assert err <= tol, f"{k}: max abs err {err} > {tol}"
    for k, ref in golden["mid"].items():
        err = (mid[k] - ref).abs().max().item()
        scale = max(1.0, ref.abs().max().item())
        assert err <= 2e-5 * scale, f"intermediate {k}: {err}"
