"""CPU: the autograd-capable restatement (oracle/torch_reference.py) against the REFERENCE'S OWN LengthRegulator
class and model modules (outputs and gradients), and against the committed goldens those produced."""
import numpy as np
import pytest
import torch

from oracle import reference_import
from oracle import torch_reference as tr
from tests import synth

needs_ref = pytest.mark.skipif(not reference_import.available(), reason="/root/reference not mounted")


def _fwd_bwd(LR, x, d, seed):
    xt = torch.from_numpy(x).requires_grad_(True)
    o, l = LR(xt, torch.from_numpy(d))
    if o.requires_grad:
        (o * torch.from_numpy(synth.upstream_grad(o.shape, seed))).sum().backward()
    g = xt.grad.numpy() if xt.grad is not None else np.zeros_like(x)
    return o.detach().numpy(), l.numpy(), g


@needs_ref
def test_restated_length_regulator_equals_reference_class_forward_and_backward():
    ref = reference_import.load()
    cases = dict(synth.lr_edge_cases())
    x, d, _ = synth.cfg2_batch(seed=2, B=4, T=60, H=16)
    cases["cfg2_small"] = (x, d)
    for name, (xe, de) in cases.items():
        a = _fwd_bwd(ref.LengthRegulator(), xe, de, 16)
        b = _fwd_bwd(tr.LengthRegulator(), xe, de, 16)
        for u, v in zip(a, b):
            assert np.array_equal(u, v), name


def test_restated_length_regulator_matches_backward_goldens(golden):
    g = golden("lr_backward.npz")
    for name, (xe, de) in synth.lr_edge_cases().items():
        _, _, grad = _fwd_bwd(tr.LengthRegulator(), xe, de, 16)
        assert np.array_equal(grad, g[f"edge_{name}_grad_x"]), name
    s = golden("lr_small.npz")
    _, _, grad = _fwd_bwd(tr.LengthRegulator(), s["x"], s["dur"], 15)
    assert np.array_equal(grad, g["small_grad_x_f32"])


def _embeddings(g):
    embs = []
    for j in range(5):
        e = torch.nn.Conv1d(1, 256, kernel_size=3, padding=1)
        with torch.no_grad():
            e.weight.copy_(torch.from_numpy(g["conv_w"][j])); e.bias.copy_(torch.from_numpy(g["conv_b"][j]))
        embs.append(e)
    return embs


def test_restated_variance_adaptor_matches_reference_model_goldens(golden):
    """forward == the reference model's own modules (variance_adaptor.npz), backward == their autograd
    (variance_adaptor_bwd.npz)."""
    g, gb = golden("variance_adaptor.npz"), golden("variance_adaptor_bwd.npz")
    embs = _embeddings(g)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    curves = [torch.from_numpy(c).requires_grad_(True) for c in g["curves"]]
    out, ml = tr.variance_adaptor(x, torch.from_numpy(g["dur"]), curves, embs)
    assert np.array_equal(ml.numpy(), g["mel_len"]) and np.array_equal(out.detach().numpy(), g["dec_input"])
    (out * torch.from_numpy(synth.upstream_grad(out.shape, seed=12))).sum().backward()
    assert np.array_equal(x.grad.numpy(), gb["grad_x"])
    assert np.array_equal(np.stack([c.grad.numpy() for c in curves]), gb["grad_curves"])
    assert np.allclose(np.stack([e.weight.grad.numpy() for e in embs]), gb["grad_w"], rtol=1e-6, atol=1e-6)
    assert np.allclose(np.stack([e.bias.grad.numpy() for e in embs]), gb["grad_b"], rtol=1e-6, atol=1e-6)
