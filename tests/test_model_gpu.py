"""GPU parity of the two TransducerGRU models (tcgen05 path) against the fp32 CPU restatement and the goldens made
from the real reference modules. Tolerances (BASELINE.json north_star): outputs within 1e-2, identical argmax on
>= 99.99 % of positions (ties closer than the tolerance are not counted as disagreements)."""
import os

import numpy as np
import pytest
import torch

import model_port as MP
from pepper_thesis_b200 import models

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-2


def _variant(seed=0):
    m = models.TransducerGRU(26, 1, 256, 28, 3, True)
    sd = MP.variant_state_dict(seed)
    m.load_state_dict(sd)
    return m, sd


def _argmax_agreement(a, b, tol):
    """fraction of rows whose argmax agrees, not counting rows where the oracle's top two are within tol"""
    aa, bb = a.argmax(-1), b.argmax(-1)
    top2 = np.sort(b, axis=-1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > tol
    return float(((aa == bb) | ~clear).mean()), float((aa == bb).mean())


def _fit_head(feat, cls, n_cls, gain):
    """Least-squares head on fp32 features: what training does to the last layer -- confident, well-separated outputs."""
    f = torch.cat([feat, torch.ones(feat.shape[0], 1)], 1).double()
    y = torch.nn.functional.one_hot(cls, n_cls).double() * 2 - 1
    sol = torch.linalg.solve(f.T @ f + 1e-3 * torch.eye(f.shape[1], dtype=torch.float64), f.T @ y)
    return (gain * sol[:-1].T).float().contiguous(), (gain * sol[-1]).float().contiguous()


def test_variant_raw_argmax_with_a_separated_head():
    """north_star: identical argmax on >= 99.99 % of positions. With random-init weights the three outputs are nearly equal
    and most disagreements are ties inside the tolerance (the other tests discount those). Here the last layer is FITTED
    (least squares on the fp32 features of three input populations) like a trained model's: outputs are confident, and the
    RAW agreement -- every row counted -- has to reach 99.99 %."""
    m, sd = _variant(6)
    g = torch.Generator().manual_seed(61)
    n = 12000
    cls = torch.arange(n) % 3
    scale = torch.tensor([6, 24, 60])[cls].view(n, 1, 1)
    x = -(torch.rand(n, 33, 26, generator=g) * scale).floor().long()
    x[:, :, 0] = torch.randint(1, 6, (n, 33), generator=g)
    feat = MP.variant_forward(sd, x.float(), return_features=True)
    sd["output_layer_type.weight"], sd["output_layer_type.bias"] = _fit_head(feat[:6000], cls[:6000], 3, 6.0)
    m.load_state_dict(sd)
    ref = MP.variant_forward(sd, x[6000:].float()).numpy()
    y = m(x[6000:], False).numpy()
    raw = float((y.argmax(-1) == ref.argmax(-1)).mean())
    margin = np.sort(ref, -1)
    print("variant, fitted head: raw argmax agreement %.5f on %d rows, max prob err %.2e, median top-2 margin %.3f, class accuracy %.3f"
          % (raw, len(ref), np.abs(y - ref).max(), float(np.median(margin[:, -1] - margin[:, -2])), float((ref.argmax(-1) == cls[6000:].numpy()).mean())))
    # the 1e-2 output tolerance is asserted on the default-initialised models (tests below). A head that separates the tiny
    # features of a random-init body needs large weights (row L1 norm ~1e4), which magnify the bf16 feature error on the few
    # rows next to a decision boundary; what this test pins is the decision itself.
    assert np.abs(y - ref).max() < 0.25
    assert raw >= 0.9999, raw


def test_polisher_raw_argmax_with_a_separated_head():
    """The same for the polisher GRU: dense1 fitted on the decoder's fp32 output of five input populations (each with its
    own pair of dominant feature columns)."""
    m, sd = _polisher(6)
    g = torch.Generator().manual_seed(62)
    n, T = 400, 100
    cls = torch.arange(n) % 5
    x = (torch.rand(n, T, 10, generator=g) * 12).floor().long()
    for i in range(n):
        x[i, :, 2 * int(cls[i])] += 60
        x[i, :, 2 * int(cls[i]) + 1] += 30
    h = torch.zeros(n, 2, 128)
    feat = MP.polisher_forward(sd, x.float(), h, return_features=True)
    pos_cls = cls.view(n, 1).expand(n, T)
    sd["dense1.weight"], sd["dense1.bias"] = _fit_head(feat[:200].reshape(-1, 256), pos_cls[:200].reshape(-1), 5, 6.0)
    m.load_state_dict(sd)
    rl, _ = MP.polisher_forward(sd, x[200:].float(), h[200:])
    logits, _ = m(x[200:], h[200:])
    raw = float((logits.numpy().argmax(-1) == rl.numpy().argmax(-1)).mean())
    top2 = np.sort(rl.numpy(), -1)
    print("polisher, fitted head: raw argmax agreement %.5f on %d positions, max logit err %.2e, median top-2 margin %.3f"
          % (raw, rl.shape[0] * rl.shape[1], (logits - rl).abs().max(), float(np.median(top2[..., -1] - top2[..., -2]))))
    assert (logits - rl).abs().max() < 0.25                   # see the variant test: large fitted weights, the decision is what is pinned
    assert raw >= 0.9999, raw


def test_variant_golden():
    g = np.load(os.path.join(GOLD, "model_variant.npz"))
    m, sd = _variant(0)
    y = m(torch.from_numpy(g["x"]), False).numpy()
    ref = g["probs"] if abs(sum(float(v.double().sum()) for v in sd.values()) - float(g["w_sum"])) < 1e-6 else \
        MP.variant_forward(sd, torch.from_numpy(g["x"].astype(np.float32))).numpy()
    assert np.abs(y - ref).max() < TOL, np.abs(y - ref).max()


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000])
def test_variant_vs_port_sizes(n):
    m, sd = _variant(1)
    g = torch.Generator().manual_seed(n)
    x = -torch.randint(0, 50, (n, 33, 26), generator=g)
    x[:, :, 0] = torch.randint(1, 6, (n, 33), generator=g)
    y = m(x, False).numpy()
    ref = MP.variant_forward(sd, x.float()).numpy()
    err = np.abs(y - ref).max()
    print("variant n=%d max prob err %.2e" % (n, err))
    assert err < TOL, err
    agree, raw = _argmax_agreement(y, ref, TOL)
    print("variant n=%d argmax agreement: raw %.5f, ties inside the tolerance discounted %.5f" % (n, raw, agree))
    assert agree >= 0.9999, (agree, raw)
    assert np.allclose(y.sum(-1), 1.0, atol=1e-5)


def test_variant_deep_coverage_and_wrap():
    """Values beyond +-127: exact hi/lo bf16 split (wrap off) and the reference pipeline's int8 round trip (wrap on)."""
    m, sd = _variant(2)
    g = torch.Generator().manual_seed(7)
    x = -torch.randint(0, 3000, (200, 33, 26), generator=g)
    y = m(x, False).numpy()
    ref = MP.variant_forward(sd, x.float()).numpy()
    assert np.abs(y - ref).max() < TOL
    probs, arg = m.infer_windows(x.to(torch.int16).cuda(), wrap_int8=True)
    xw = x.to(torch.int16).numpy().astype(np.int8).astype(np.float32)      # DataStore.py:68 / dataloader_predict.py:90
    refw = MP.variant_forward(sd, torch.from_numpy(xw)).numpy()
    assert np.abs(probs.cpu().numpy() - refw).max() < TOL
    assert (arg.cpu().numpy() == probs.cpu().numpy().argmax(-1)).all()


def test_variant_large_batch_chunks():
    m, sd = _variant(3)
    g = torch.Generator().manual_seed(11)
    n = 9000                                                   # > one 8192-window pass
    x = -torch.randint(0, 40, (n, 33, 26), generator=g)
    y = m(x, False).numpy()
    idx = np.r_[0:64, 8150:8250, n - 64:n]
    ref = MP.variant_forward(sd, x[idx].float()).numpy()
    assert np.abs(y[idx] - ref).max() < TOL


def test_variant_checkpoint_contract(tmp_path):
    sd = MP.variant_state_dict(5)
    ck = {"model_state_dict": {"module." + k: v for k, v in sd.items()}, "model_optimizer": {}, "hidden_size": 256,
          "gru_layers": 1, "epochs": 3}
    path = str(tmp_path / "ckpt.pkl")
    torch.save(ck, path)
    m, hs, gl, ep = models.ModelHandler.load_simple_model_for_training(path, 26, 28, 3)
    assert (hs, gl, ep) == (256, 1, 3)
    x = -torch.randint(0, 30, (16, 33, 26))
    assert np.abs(m(x, False).numpy() - MP.variant_forward(sd, x.float()).numpy()).max() < TOL


def _polisher(seed=0):
    m = models.PolisherTransducerGRU(1, 10, 1, 128, 5, True)
    sd = MP.polisher_state_dict(seed)
    m.load_state_dict(sd)
    return m, sd


def test_polisher_golden():
    g = np.load(os.path.join(GOLD, "model_polisher.npz"))
    m, sd = _polisher(0)
    logits, hf = m(torch.from_numpy(g["x"]), torch.from_numpy(g["h"]))
    if abs(sum(float(v.double().sum()) for v in sd.values()) - float(g["w_sum"])) < 1e-6:
        rl, rh = g["logits"], g["h_final"]
    else:
        a, b = MP.polisher_forward(sd, torch.from_numpy(g["x"].astype(np.float32)), torch.from_numpy(g["h"]))
        rl, rh = a.numpy(), b.numpy()
    assert np.abs(logits.numpy() - rl).max() < TOL, np.abs(logits.numpy() - rl).max()
    assert np.abs(hf.numpy() - rh).max() < TOL


@pytest.mark.parametrize("n", [3, 130, 600])
def test_polisher_vs_port(n):
    m, sd = _polisher(1)
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 255, (n, 100, 10), generator=g)
    h = torch.randn(n, 2, 128, generator=g) * 0.5
    logits, hf = m(x, h)
    rl, rh = MP.polisher_forward(sd, x.float(), h)
    print("polisher n=%d max logit err %.2e hidden err %.2e" % (n, (logits - rl).abs().max(), (hf - rh).abs().max()))
    assert (logits - rl).abs().max() < TOL, (logits - rl).abs().max()
    assert (hf - rh).abs().max() < TOL
    agree, raw = _argmax_agreement(logits.numpy(), rl.numpy(), TOL)
    print("polisher n=%d argmax agreement: raw %.5f, ties inside the tolerance discounted %.5f" % (n, raw, agree))
    assert agree >= 0.9999, (agree, raw)


@pytest.mark.parametrize("n,T", [(1, 100), (128, 1), (129, 7), (257, 33)])
def test_polisher_ragged_rows_and_lengths(n, T):
    """Row counts around the 128-window tile of the recurrence kernel, sequence lengths other than the window of 100."""
    m, sd = _polisher(3)
    g = torch.Generator().manual_seed(100 * n + T)
    x = torch.randint(0, 255, (n, T, 10), generator=g)
    h = torch.randn(n, 2, 128, generator=g) * 0.5
    logits, hf = m(x, h)
    rl, rh = MP.polisher_forward(sd, x.float(), h)
    assert logits.shape == rl.shape and hf.shape == rh.shape
    assert (logits - rl).abs().max() < TOL, (logits - rl).abs().max()
    assert (hf - rh).abs().max() < TOL, (hf - rh).abs().max()


def test_polisher_more_windows_than_one_pass():
    """More windows than one pass of the model holds (16384): the second pass reuses the gx / layer-output workspaces."""
    m, sd = _polisher(4)
    n, T = 16384 + 130, 3
    g = torch.Generator().manual_seed(9)
    x = torch.randint(0, 255, (n, T, 10), generator=g)
    h = torch.randn(n, 2, 128, generator=g) * 0.5
    logits, hf = m(x, h)
    rl, rh = MP.polisher_forward(sd, x.float(), h)
    assert (logits - rl).abs().max() < TOL, (logits - rl).abs().max()
    assert (hf - rh).abs().max() < TOL, (hf - rh).abs().max()


def test_polisher_chunk_loop():
    """1000-position chunks, 100-wide windows every 50, hidden carried, softmax summed (predict_distributed_gpu.py:63-96)."""
    m, sd = _polisher(2)
    g = torch.Generator().manual_seed(5)
    x = torch.randint(0, 255, (20, 1000, 10), generator=g)
    acc, labels = m.predict_chunks(x, 100, 50)
    racc, rlab = MP.polisher_predict_chunks(sd, x, 100, 50)
    assert (acc.cpu() - racc).abs().max() < 2 * TOL
    agree, raw = _argmax_agreement(acc.cpu().numpy(), racc.numpy(), 2 * TOL)
    assert agree >= 0.9999, (agree, raw)
    assert (labels.cpu().numpy() == acc.cpu().numpy().argmax(-1)).all()


def test_polisher_folded_input_projection_variant(monkeypatch):
    """PV_GRU_FOLD=1: layer 1 without the gx round trip (r / z input parts on the MMA through a third k atom, the candidate's on
    fp32 FMAs in the epilogue, x_t streamed 16 bytes per window and step): same outputs as the default path within the
    tolerance both have against the fp32 port."""
    monkeypatch.setenv("PV_GRU_FOLD", "1")
    m, sd = _polisher(1)                                        # the variable is read when the model is created
    monkeypatch.delenv("PV_GRU_FOLD")
    m0, _ = _polisher(1)
    for n, T in ((300, 100), (129, 37), (5, 100)):
        g = torch.Generator().manual_seed(7 * n + T)
        x = torch.randint(0, 255, (n, T, 10), generator=g)
        h = torch.randn(n, 2, 128, generator=g) * 0.5
        la, ha = m(x, h.clone())
        lb, hb = m0(x, h.clone())
        rl, rh = MP.polisher_forward(sd, x.float(), h)
        assert (la - rl).abs().max() < TOL and (ha - rh).abs().max() < TOL
        assert (la - lb).abs().max() < TOL and (ha - hb).abs().max() < TOL
