"""The depth-encoder oracle (oracle/depth_encoder_oracle.py) against golden vectors generated from the unmodified
reference module (oracle/make_golden_encoder.py -> tests/golden/depth_encoder.npz)."""
import os

import numpy as np
import torch

from oracle import depth_encoder_oracle as EO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "depth_encoder.npz")


def test_encoder_oracle_matches_reference_golden():
    rec = np.load(GOLD)
    sd = EO.make_weights(700, 703)
    sums = np.array([float(sd[k].double().sum()) for k in EO.KEYS])
    assert np.allclose(sums, rec["sd0_sums"], rtol=0, atol=1e-9), "initial weights differ from the reference's"
    for k in EO.KEYS:
        if "sd0." + k in rec.files:
            assert np.array_equal(sd[k].numpy(), rec["sd0." + k]), k
    params = {k: sd[k].clone().requires_grad_(True) for k in EO.KEYS}
    state = dict(sd)
    state.update(params)
    x = EO.make_inputs(2, 701)
    feats = EO.encoder_forward(state, x, training=True)
    assert feats.shape == (2, 196, 2048)
    assert np.allclose(feats.detach()[:, ::7, ::64].numpy(), rec["train_sub"], rtol=1e-5, atol=1e-6)
    tot = np.array([float(feats.detach().double().sum()), float(feats.detach().double().abs().sum())])
    assert np.allclose(tot, rec["train_sum"], rtol=1e-6)
    (feats * EO.projection(feats.shape, 702)).sum().backward()
    for k in EO.KEYS:
        g = params[k].grad
        ref_norm = float(rec["gnorm." + k][0])
        assert abs(float(g.double().norm()) - ref_norm) <= 1e-4 * ref_norm + 1e-6, k
        assert np.allclose(g.flatten()[:64].numpy(), rec["ghead." + k], rtol=2e-3, atol=2e-4 * ref_norm), k
    for i in (1, 2, 3):
        assert np.allclose(state[f"bn{i}.running_mean"].numpy(), rec[f"rm{i}"], rtol=1e-5, atol=1e-7)
        assert np.allclose(state[f"bn{i}.running_var"].numpy(), rec[f"rv{i}"], rtol=1e-5, atol=1e-7)
    with torch.no_grad():
        fe = EO.encoder_forward(state, x, training=False)
    assert np.allclose(fe[:, ::7, ::64].numpy(), rec["eval_sub"], rtol=1e-5, atol=1e-6)
    # the 14 x 14 map is the exact 2 x 2 replication of a 7 x 7 one (what the CUDA path builds on)
    f4 = feats.detach().reshape(2, 7, 2, 7, 2, 2048)
    assert torch.equal(f4[:, :, 0, :, 0], f4[:, :, 1, :, 1]) and torch.equal(f4[:, :, 0, :, 0], f4[:, :, 0, :, 1])
