"""TransducerGRU model contracts on the CUDA path.

``TransducerGRU`` (variant, model M-A) mirrors /root/reference/pepper_variant/modules/python/models/simple_model.py:6-82:
same constructor arguments, same ``state_dict`` keys/shapes, ``model(images, False) -> softmax [B, 3]``
(predict_distributed_gpu.py:65). ``PolisherTransducerGRU`` (model M-B) mirrors
/root/reference/pepper/modules/python/models/simple_model.py:5-49: ``model(x, hidden) -> (logits, hidden_final)``.
``ModelHandler`` mirrors ModelHander.py:18-44 (checkpoint dict with ``model_state_dict`` / ``hidden_size`` /
``gru_layers`` / ``epochs``, optional ``module.`` prefix).

There is no eager/PyTorch fallback: the forward pass is ``pv_lstm_infer`` / ``pv_gru_forward`` in libpepper_b200.so
and raises ``PvError`` when the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import capi

VARIANT_SHAPES = OrderedDict(
    [("encoder.weight_ih_l0", (1024, 26)), ("encoder.weight_hh_l0", (1024, 256)), ("encoder.bias_ih_l0", (1024,)),
     ("encoder.bias_hh_l0", (1024,)), ("encoder.weight_ih_l0_reverse", (1024, 26)),
     ("encoder.weight_hh_l0_reverse", (1024, 256)), ("encoder.bias_ih_l0_reverse", (1024,)),
     ("encoder.bias_hh_l0_reverse", (1024,)), ("decoder.weight_ih_l0", (1024, 512)),
     ("decoder.weight_hh_l0", (1024, 256)), ("decoder.bias_ih_l0", (1024,)), ("decoder.bias_hh_l0", (1024,)),
     ("decoder.weight_ih_l0_reverse", (1024, 512)), ("decoder.weight_hh_l0_reverse", (1024, 256)),
     ("decoder.bias_ih_l0_reverse", (1024,)), ("decoder.bias_hh_l0_reverse", (1024,)),
     ("linear_1.weight", (512, 16896)), ("linear_1.bias", (512,)), ("linear_2.weight", (512, 512)),
     ("linear_2.bias", (512,)), ("linear_3.weight", (512, 512)), ("linear_3.bias", (512,)),
     ("linear_4.weight", (512, 512)), ("linear_4.bias", (512,)), ("linear_5.weight", (512, 512)),
     ("linear_5.bias", (512,)), ("output_layer_type.weight", (3, 512)), ("output_layer_type.bias", (3,))])

POLISHER_SHAPES = OrderedDict(
    [("gru_encoder.weight_ih_l0", (384, 10)), ("gru_encoder.weight_hh_l0", (384, 128)),
     ("gru_encoder.bias_ih_l0", (384,)), ("gru_encoder.bias_hh_l0", (384,)),
     ("gru_encoder.weight_ih_l0_reverse", (384, 10)), ("gru_encoder.weight_hh_l0_reverse", (384, 128)),
     ("gru_encoder.bias_ih_l0_reverse", (384,)), ("gru_encoder.bias_hh_l0_reverse", (384,)),
     ("gru_decoder.weight_ih_l0", (384, 256)), ("gru_decoder.weight_hh_l0", (384, 128)),
     ("gru_decoder.bias_ih_l0", (384,)), ("gru_decoder.bias_hh_l0", (384,)),
     ("gru_decoder.weight_ih_l0_reverse", (384, 256)), ("gru_decoder.weight_hh_l0_reverse", (384, 128)),
     ("gru_decoder.bias_ih_l0_reverse", (384,)), ("gru_decoder.bias_hh_l0_reverse", (384,)),
     ("dense1.weight", (5, 256)), ("dense1.bias", (5,))])


def _check_state_dict(sd, shapes):
    clean = OrderedDict()
    for k, v in sd.items():
        clean[k[7:] if k.startswith("module.") else k] = v           # ModelHander.py:34-40
    missing = [k for k in shapes if k not in clean]
    if missing:
        raise KeyError("missing keys in state_dict: %s" % missing)
    out = OrderedDict()
    for k, shp in shapes.items():
        t = torch.as_tensor(clean[k]).detach().to("cpu", torch.float32).contiguous()
        if tuple(t.shape) != tuple(shp):
            raise ValueError("size mismatch for %s: %s vs %s" % (k, tuple(t.shape), shp))
        out[k] = t
    return out


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class TransducerGRU:
    """Variant model M-A (2x biLSTM-256 + SELU MLP + softmax over 3 genotype classes)."""

    def __init__(self, image_features=26, gru_layers=1, hidden_size=256, num_classes=28, num_classes_type=3,
                 bidirectional=True):
        if image_features != 26 or gru_layers != 1 or not bidirectional or num_classes_type != 3:
            raise ValueError("the CUDA path is built for image_features=26, gru_layers=1, bidirectional, 3 type classes")
        self.hidden_size = hidden_size
        self.num_layers = gru_layers
        self.num_classes = num_classes
        self.num_classes_type = num_classes_type
        self.wrap_int8 = True        # reproduce the reference pipeline's int8 HDF5 round trip (DataStore.py:68)
        self._sd = None
        self._handle = None
        self._ws = None

    # -- nn.Module-like surface ----------------------------------------------------------------------------------
    def state_dict(self):
        return OrderedDict((k, v.clone()) for k, v in self._sd.items())

    def load_state_dict(self, sd):
        self._sd = _check_state_dict(sd, VARIANT_SHAPES)
        self._destroy()
        return self

    def eval(self):
        return self

    def cuda(self, device=None):
        return self

    def cpu(self):
        return self

    def _destroy(self):
        if self._handle is not None:
            capi.load().pv_lstm_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _ensure(self):
        if self._handle is not None:
            return
        if self._sd is None:
            raise RuntimeError("load_state_dict first (random init: oracle-free helper models.random_variant_state_dict)")
        lib = capi.load()
        sd, w = self._sd, capi.PvLstmWeightsStruct()
        for d, suf in enumerate(("", "_reverse")):
            w.enc_w_ih[d] = sd["encoder.weight_ih_l0" + suf].data_ptr()
            w.enc_w_hh[d] = sd["encoder.weight_hh_l0" + suf].data_ptr()
            w.enc_b_ih[d] = sd["encoder.bias_ih_l0" + suf].data_ptr()
            w.enc_b_hh[d] = sd["encoder.bias_hh_l0" + suf].data_ptr()
            w.dec_w_ih[d] = sd["decoder.weight_ih_l0" + suf].data_ptr()
            w.dec_w_hh[d] = sd["decoder.weight_hh_l0" + suf].data_ptr()
            w.dec_b_ih[d] = sd["decoder.bias_ih_l0" + suf].data_ptr()
            w.dec_b_hh[d] = sd["decoder.bias_hh_l0" + suf].data_ptr()
        for l in range(5):
            w.lin_w[l] = sd["linear_%d.weight" % (l + 1)].data_ptr()
            w.lin_b[l] = sd["linear_%d.bias" % (l + 1)].data_ptr()
        w.out_w = sd["output_layer_type.weight"].data_ptr()
        w.out_b = sd["output_layer_type.bias"].data_ptr()
        h = C.c_void_p()
        capi.check(lib.pv_lstm_create(C.byref(w), C.byref(h)))
        self._handle = h

    def infer_windows(self, windows: torch.Tensor, wrap_int8=None):
        """windows int16 [n,33,26] on the GPU -> (probs float32 [n,3], argmax uint8 [n]) on the GPU, current stream."""
        self._ensure()
        lib = capi.load()
        assert windows.is_cuda and windows.dtype == torch.int16 and windows.is_contiguous()
        n = int(windows.shape[0])
        probs = torch.empty((n, 3), dtype=torch.float32, device=windows.device)
        arg = torch.empty(n, dtype=torch.uint8, device=windows.device)
        if n == 0:
            return probs, arg
        need = int(lib.pv_lstm_workspace_bytes(n))
        if self._ws is None or self._ws.numel() < need or self._ws.device != windows.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=windows.device)
        wrap = self.wrap_int8 if wrap_int8 is None else wrap_int8
        stream = torch.cuda.current_stream(windows.device).cuda_stream
        capi.check(lib.pv_lstm_infer(self._handle, _ptr(windows), n, int(bool(wrap)), _ptr(probs), _ptr(arg),
                                     _ptr(self._ws), self._ws.numel(), C.c_void_p(stream)))
        return probs, arg

    def forward(self, x, train_mode=False):
        """``model(images, False)`` of predict_distributed_gpu.py:65: x [B,33,26] of integer-valued counts (any
        dtype/device) -> softmax [B,3] float32 on x's device."""
        if train_mode:
            raise NotImplementedError("inference only")
        if x.dim() != 3 or x.shape[1] != 33 or x.shape[2] != 26:
            raise ValueError("expected [B, 33, 26]")
        xi = x.to(torch.int16) if x.dtype != torch.int16 else x
        dev = x.device if x.is_cuda else torch.device("cuda")
        probs, _ = self.infer_windows(xi.to(dev).contiguous(), wrap_int8=False)
        return probs.to(x.device)

    __call__ = forward


class PolisherTransducerGRU:
    """Polisher model M-B (2x biGRU-128 + Linear(256, 5)); ``model(x, hidden) -> (logits, hidden_final)``."""

    def __init__(self, image_channels=1, image_features=10, gru_layers=1, hidden_size=128, num_classes=5, bidirectional=True):
        if image_features != 10 or gru_layers != 1 or hidden_size != 128 or num_classes != 5 or not bidirectional:
            raise ValueError("the CUDA path is built for image_features=10, hidden 128, 5 classes, bidirectional")
        self.hidden_size = hidden_size
        self.num_layers = gru_layers
        self.num_classes = num_classes
        self._sd = None
        self._handle = None
        self._ws = None

    def state_dict(self):
        return OrderedDict((k, v.clone()) for k, v in self._sd.items())

    def load_state_dict(self, sd):
        self._sd = _check_state_dict(sd, POLISHER_SHAPES)
        self._destroy()
        return self

    def eval(self):
        return self

    def init_hidden(self, batch_size, num_layers=1, bidirectional=True):
        return torch.zeros(batch_size, 2 * num_layers, self.hidden_size)

    def _destroy(self):
        if self._handle is not None:
            capi.load().pv_gru_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _ensure(self):
        if self._handle is not None:
            return
        lib = capi.load()
        sd, w = self._sd, capi.PvGruWeightsStruct()
        for d, suf in enumerate(("", "_reverse")):
            w.enc_w_ih[d] = sd["gru_encoder.weight_ih_l0" + suf].data_ptr()
            w.enc_w_hh[d] = sd["gru_encoder.weight_hh_l0" + suf].data_ptr()
            w.enc_b_ih[d] = sd["gru_encoder.bias_ih_l0" + suf].data_ptr()
            w.enc_b_hh[d] = sd["gru_encoder.bias_hh_l0" + suf].data_ptr()
            w.dec_w_ih[d] = sd["gru_decoder.weight_ih_l0" + suf].data_ptr()
            w.dec_w_hh[d] = sd["gru_decoder.weight_hh_l0" + suf].data_ptr()
            w.dec_b_ih[d] = sd["gru_decoder.bias_ih_l0" + suf].data_ptr()
            w.dec_b_hh[d] = sd["gru_decoder.bias_hh_l0" + suf].data_ptr()
        w.dense_w = sd["dense1.weight"].data_ptr()
        w.dense_b = sd["dense1.bias"].data_ptr()
        h = C.c_void_p()
        capi.check(lib.pv_gru_create(C.byref(w), C.byref(h)))
        self._handle = h

    def _workspace(self, n, seq_len, device):
        need = int(capi.load().pv_gru_workspace_bytes(n, seq_len))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def forward(self, x, hidden):
        """x [B,T,10] integer-valued 0..254 counts, hidden [B,2,128] -> (logits [B,T,5], hidden_final [B,2,128])."""
        self._ensure()
        lib = capi.load()
        dev = x.device if x.is_cuda else torch.device("cuda")
        xi = x.to(dev).to(torch.uint8).contiguous()
        h = hidden.to(dev, torch.float32).contiguous().clone()
        n, t = int(xi.shape[0]), int(xi.shape[1])
        logits = torch.empty((n, t, 5), dtype=torch.float32, device=dev)
        ws = self._workspace(n, t, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        capi.check(lib.pv_gru_forward(self._handle, _ptr(xi), n, t, _ptr(h), _ptr(logits), _ptr(ws), ws.numel(),
                                      C.c_void_p(stream)))
        return logits.to(x.device), h.to(hidden.device)

    __call__ = forward

    def predict_chunks(self, images, window=100, stride=50):
        """The polisher's chunk loop (predict_distributed_gpu.py:63-96): images uint8 [B,L,10] ->
        (summed softmax [B,L,5], labels uint8 [B,L])."""
        self._ensure()
        lib = capi.load()
        dev = images.device if images.is_cuda else torch.device("cuda")
        xi = images.to(dev).to(torch.uint8).contiguous()
        n, L = int(xi.shape[0]), int(xi.shape[1])
        acc = torch.zeros((n, L, 5), dtype=torch.float32, device=dev)
        labels = torch.empty((n, L), dtype=torch.uint8, device=dev)
        ws = self._workspace(n, window, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        capi.check(lib.pv_gru_predict_chunks(self._handle, _ptr(xi), n, L, window, stride, _ptr(acc), _ptr(labels),
                                             _ptr(ws), ws.numel(), C.c_void_p(stream)))
        return acc, labels


def random_variant_state_dict(seed=0):
    """Random-init weights with the reference's parameter creation order (simple_model.py:23-46)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    enc = nn.LSTM(26, 256, num_layers=1, bidirectional=True, batch_first=True)
    dec = nn.LSTM(512, 256, num_layers=1, bidirectional=True, batch_first=True)
    lins = [nn.Linear(16896, 512)] + [nn.Linear(512, 512) for _ in range(4)]
    out = nn.Linear(512, 3)
    sd = OrderedDict()
    for name, mod in (("encoder", enc), ("decoder", dec)):
        for k, v in mod.state_dict().items():
            sd[name + "." + k] = v.detach().clone()
    for i, l in enumerate(lins):
        sd["linear_%d.weight" % (i + 1)] = l.weight.detach().clone()
        sd["linear_%d.bias" % (i + 1)] = l.bias.detach().clone()
    sd["output_layer_type.weight"] = out.weight.detach().clone()
    sd["output_layer_type.bias"] = out.bias.detach().clone()
    return sd


def random_polisher_state_dict(seed=0):
    import torch.nn as nn
    torch.manual_seed(seed)
    enc = nn.GRU(10, 128, num_layers=1, bidirectional=True, batch_first=True)
    dec = nn.GRU(256, 128, num_layers=1, bidirectional=True, batch_first=True)
    dense = nn.Linear(256, 5)
    sd = OrderedDict()
    for name, mod in (("gru_encoder", enc), ("gru_decoder", dec)):
        for k, v in mod.state_dict().items():
            sd[name + "." + k] = v.detach().clone()
    sd["dense1.weight"] = dense.weight.detach().clone()
    sd["dense1.bias"] = dense.bias.detach().clone()
    return sd


class ModelHandler:
    """ModelHander.py:5-44 (variant)."""

    @staticmethod
    def save_checkpoint(state, filename):
        torch.save(state, filename)

    @staticmethod
    def get_new_gru_model(image_features, gru_layers, hidden_size, num_classes, num_classes_type):
        return TransducerGRU(image_features, gru_layers, hidden_size, num_classes, num_classes_type, bidirectional=True)

    @staticmethod
    def load_simple_model_for_training(model_path, image_features, num_classes, num_type_classes):
        checkpoint = torch.load(model_path, map_location="cpu")
        hidden_size, gru_layers, epochs = checkpoint["hidden_size"], checkpoint["gru_layers"], checkpoint["epochs"]
        model = ModelHandler.get_new_gru_model(image_features, gru_layers, hidden_size, num_classes, num_type_classes)
        model.load_state_dict(checkpoint["model_state_dict"])
        return model, hidden_size, gru_layers, epochs
