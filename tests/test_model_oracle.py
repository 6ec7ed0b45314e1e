"""CPU: the model restatement (oracle/model_port.py) against the golden vectors made from the real reference modules."""
import os

import numpy as np
import pytest
import torch

import model_port as MP

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _wsum(sd):
    return sum(float(v.double().sum()) for v in sd.values())


def test_variant_port_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "model_variant.npz"))
    sd = MP.variant_state_dict(0)
    if abs(_wsum(sd) - float(g["w_sum"])) > 1e-6:
        pytest.skip("torch default init differs from the build container's (different torch build)")
    y = MP.variant_forward(sd, torch.from_numpy(g["x"].astype(np.float32)))
    assert np.abs(y.numpy() - g["probs"]).max() < 1e-5


def test_polisher_port_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "model_polisher.npz"))
    sd = MP.polisher_state_dict(0)
    if abs(_wsum(sd) - float(g["w_sum"])) > 1e-6:
        pytest.skip("torch default init differs from the build container's (different torch build)")
    logits, hf = MP.polisher_forward(sd, torch.from_numpy(g["x"].astype(np.float32)), torch.from_numpy(g["h"]))
    assert np.abs(logits.numpy() - g["logits"]).max() < 1e-4
    assert np.abs(hf.numpy() - g["h_final"]).max() < 1e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/pepper_variant"), reason="reference tree absent")
def test_ports_match_live_reference():
    import sys
    sys.path.insert(0, "/root/reference")
    from pepper_variant.modules.python.models.simple_model import TransducerGRU as V
    from pepper.modules.python.models.simple_model import TransducerGRU as P
    torch.manual_seed(3)
    m = V(26, 1, 256, 28, 3, True).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x = torch.randint(-50, 6, (5, 33, 26)).float()
    with torch.no_grad():
        assert (m(x, False) - MP.variant_forward(sd, x)).abs().max() < 1e-6
    p = P(1, 10, 1, 128, 5, True).eval()
    sd = {k: v.detach() for k, v in p.state_dict().items()}
    x = torch.randint(0, 255, (3, 100, 10)).float()
    h = torch.randn(3, 2, 128)
    with torch.no_grad():
        a, ah = p(x, h)
    b, bh = MP.polisher_forward(sd, x, h)
    assert (a - b).abs().max() < 1e-4 and (ah - bh).abs().max() < 1e-5
