"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two TransducerGRU models (fp32, plain torch tensor ops).

The reference's arithmetic for this part lives in PyTorch (torch==1.10.0 pinned in /root/reference/requirements.txt:3;
this image has torch 2.11): nn.LSTM / nn.GRU / nn.Linear / nn.SELU / nn.Softmax. This file restates the module wiring
and the recurrences explicitly (no nn.LSTM / nn.GRU call), following
  variant  /root/reference/pepper_variant/modules/python/models/simple_model.py:6-82
  polisher /root/reference/pepper/modules/python/models/simple_model.py:5-42 and the chunk loop
           /root/reference/pepper/modules/python/models/predict_distributed_gpu.py:63-96
Parity status: pinned against the imported reference classes in the build container (tests/golden/make_model_golden.py
writes tests/golden/model_*.npz from the REAL reference modules; tests/test_model_oracle.py checks this port against
those files everywhere and against the live reference when /root/reference exists).
"""
import torch
import torch.nn as nn


def variant_state_dict(seed=0, image_features=26, hidden=256, classes_type=3, window=33):
    """Default-initialised weights with the parameter creation order of simple_model.py:23-46."""
    torch.manual_seed(seed)
    enc = nn.LSTM(image_features, hidden, num_layers=1, bidirectional=True, batch_first=True)
    dec = nn.LSTM(2 * hidden, hidden, num_layers=1, bidirectional=True, batch_first=True)
    lins = [nn.Linear(2 * hidden * window, 512)] + [nn.Linear(512, 512) for _ in range(4)]
    out = nn.Linear(512, classes_type)
    sd = {}
    for name, mod in (("encoder", enc), ("decoder", dec)):
        for k, v in mod.state_dict().items():
            sd[name + "." + k] = v.detach().clone()
    for i, l in enumerate(lins):
        sd["linear_%d.weight" % (i + 1)] = l.weight.detach().clone()
        sd["linear_%d.bias" % (i + 1)] = l.bias.detach().clone()
    sd["output_layer_type.weight"] = out.weight.detach().clone()
    sd["output_layer_type.bias"] = out.bias.detach().clone()
    return sd


def polisher_state_dict(seed=0, image_features=10, hidden=128, classes=5):
    """Creation order of pepper/modules/python/models/simple_model.py:12-24."""
    torch.manual_seed(seed)
    enc = nn.GRU(image_features, hidden, num_layers=1, bidirectional=True, batch_first=True)
    dec = nn.GRU(2 * hidden, hidden, num_layers=1, bidirectional=True, batch_first=True)
    dense = nn.Linear(2 * hidden, classes)
    sd = {}
    for name, mod in (("gru_encoder", enc), ("gru_decoder", dec)):
        for k, v in mod.state_dict().items():
            sd[name + "." + k] = v.detach().clone()
    sd["dense1.weight"] = dense.weight.detach().clone()
    sd["dense1.bias"] = dense.bias.detach().clone()
    return sd


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM, batch_first, zero initial state; gate rows i,f,g,o."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H); c = x.new_zeros(B, H)
    out = x.new_zeros(B, T, H)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        g = x[:, t] @ w_ih.T + b_ih + h @ w_hh.T + b_hh
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def _bilstm(x, sd, name):
    f = _lstm_dir(x, sd[name + ".weight_ih_l0"], sd[name + ".weight_hh_l0"], sd[name + ".bias_ih_l0"], sd[name + ".bias_hh_l0"], False)
    r = _lstm_dir(x, sd[name + ".weight_ih_l0_reverse"], sd[name + ".weight_hh_l0_reverse"], sd[name + ".bias_ih_l0_reverse"],
                  sd[name + ".bias_hh_l0_reverse"], True)
    return torch.cat([f, r], dim=2)


def variant_forward(sd, x, return_logits=False, return_features=False):
    """simple_model.py:48-82 in eval mode (dropout = identity). x float32 [B,33,26] -> softmax [B,3]."""
    with torch.no_grad():
        x = x.float()
        x = _bilstm(x, sd, "encoder")
        x = _bilstm(x, sd, "decoder")
        x = torch.flatten(x, 1, 2)
        for i in range(1, 6):
            x = torch.selu(x @ sd["linear_%d.weight" % i].T + sd["linear_%d.bias" % i])
        if return_features:                                   # input of output_layer_type (tests fit a separated head on it)
            return x
        logits = x @ sd["output_layer_type.weight"].T + sd["output_layer_type.bias"]
        return logits if return_logits else torch.softmax(logits, dim=1)


def _gru_dir(x, h0, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.GRU; gate rows r,z,n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn))."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = h0
    out = x.new_zeros(B, T, H)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        gi = x[:, t] @ w_ih.T + b_ih
        gh = h @ w_hh.T + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        out[:, t] = h
    return out, h


def _bigru(x, h0, sd, name):
    f, hf = _gru_dir(x, h0[0], sd[name + ".weight_ih_l0"], sd[name + ".weight_hh_l0"], sd[name + ".bias_ih_l0"], sd[name + ".bias_hh_l0"], False)
    r, hr = _gru_dir(x, h0[1], sd[name + ".weight_ih_l0_reverse"], sd[name + ".weight_hh_l0_reverse"], sd[name + ".bias_ih_l0_reverse"],
                     sd[name + ".bias_hh_l0_reverse"], True)
    return torch.cat([f, r], dim=2), torch.stack([hf, hr], dim=0)


def polisher_forward(sd, x, hidden, return_features=False):
    """pepper simple_model.py:27-42. x [B,T,10] float, hidden [B,2,128] -> logits [B,T,5], hidden_final [B,2,128]."""
    with torch.no_grad():
        h0 = hidden.float().transpose(0, 1).contiguous()
        x_out, h_enc = _bigru(x.float(), h0, sd, "gru_encoder")
        x_out, h_dec = _bigru(x_out, h_enc, sd, "gru_decoder")            # decoder h0 = encoder h_n
        if return_features:                                   # input of dense1
            return x_out
        logits = x_out @ sd["dense1.weight"].T + sd["dense1.bias"]
        return logits, h_dec.transpose(0, 1).contiguous()


def polisher_predict_chunks(sd, images, window=100, stride=50):
    """predict_distributed_gpu.py:63-96: sliding windows, carried hidden, summed softmax, argmax."""
    B, L, _ = images.shape
    hidden = torch.zeros(B, 2, sd["gru_encoder.weight_hh_l0"].shape[1])
    acc = torch.zeros(B, L, sd["dense1.weight"].shape[0])
    for i in range(0, L, stride):
        if i + window > L:
            break
        logits, hidden = polisher_forward(sd, images[:, i:i + window].float(), hidden)
        acc[:, i:i + window] += torch.softmax(logits, dim=2)
    return acc, acc.argmax(dim=2)


class TorchVariantModule(nn.Module):
    """nn.Module re-declaration of the variant TransducerGRU (simple_model.py:6-82) on the stock PyTorch operators
    (nn.LSTM / nn.Linear / nn.SELU / nn.Softmax) -- used ONLY as the CPU baseline of bench.py, where the reference's
    own module cannot be imported (its tree is absent on the GPU box)."""

    def __init__(self, sd):
        super().__init__()
        self.encoder = nn.LSTM(26, 256, num_layers=1, bidirectional=True, batch_first=True)
        self.decoder = nn.LSTM(512, 256, num_layers=1, bidirectional=True, batch_first=True)
        self.activation = nn.SELU()
        self.linear_1 = nn.Linear(512 * 33, 512)
        self.linear_2 = nn.Linear(512, 512)
        self.linear_3 = nn.Linear(512, 512)
        self.linear_4 = nn.Linear(512, 512)
        self.linear_5 = nn.Linear(512, 512)
        self.output_layer_type = nn.Linear(512, 3)
        self.load_state_dict(sd)

    def forward(self, x):
        x, _ = self.encoder(x)
        x, _ = self.decoder(x)
        x = torch.flatten(x, start_dim=1, end_dim=2)
        for lin in (self.linear_1, self.linear_2, self.linear_3, self.linear_4, self.linear_5):
            x = self.activation(lin(x))
        return torch.softmax(self.output_layer_type(x), dim=1)
