"""Writes tests/golden/model_variant.npz and model_polisher.npz from the REAL reference modules
(/root/reference/pepper_variant/.../simple_model.py and /root/reference/pepper/.../simple_model.py), imported in the
build container. Weights are not stored: they are torch.manual_seed(0) default initialisations (the same torch build
runs on the GPU box); inputs and reference outputs are stored.
    python tests/golden/make_model_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from pepper_variant.modules.python.models.simple_model import TransducerGRU as Variant  # noqa: E402
from pepper.modules.python.models.simple_model import TransducerGRU as Polisher  # noqa: E402

torch.manual_seed(0)
m = Variant(26, 1, 256, 28, 3, True).eval()
g = torch.Generator().manual_seed(1234)
x = -torch.randint(0, 60, (64, 33, 26), generator=g)
x[:, :, 0] = torch.randint(1, 6, (64, 33), generator=g)
x[:, :, 1:4] = 0
x[:, 16, 5] = torch.randint(0, 30, (64,), generator=g)
x[:8] *= 3                      # some rows beyond the int8 / bf16-exact range (|x| up to 177)
with torch.no_grad():
    y = m(x.float(), False)
np.savez_compressed(os.path.join(HERE, "model_variant.npz"), x=x.numpy().astype(np.int16), probs=y.numpy(),
                    w_sum=np.float64(sum(float(v.double().sum()) for v in m.state_dict().values())))

torch.manual_seed(0)
p = Polisher(1, 10, 1, 128, 5, True).eval()
xi = torch.randint(0, 255, (16, 100, 10), generator=g)
h = torch.randn(16, 2, 128, generator=g) * 0.3
with torch.no_grad():
    logits, hf = p(xi.float(), h)
np.savez_compressed(os.path.join(HERE, "model_polisher.npz"), x=xi.numpy().astype(np.uint8), h=h.numpy(),
                    logits=logits.numpy(), h_final=hf.numpy(),
                    w_sum=np.float64(sum(float(v.double().sum()) for v in p.state_dict().values())))
print("ok", y[:2], logits[0, 0])
