"""The training-step oracle (oracle.model_gradients / adam_step) pinned against an independent implementation:
torch autograd + torch.optim.Adam (weight_decay = L2 added to the gradient, like dfdx's WeightDecay::L2).
The reference has no test for update_model and dfdx is not vendored: parity unpinned by the
reference; this pins the restatement of nabla/model/dfdx.rs:86-131 to PyTorch fp32 instead."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _torch_model(params, dims):
    layers, off = [], 0
    for l in range(4):
        din, dout = dims[l], dims[l + 1]
        lin = torch.nn.Linear(din, dout)
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(params[off:off + din * dout].reshape(dout, din).copy()))
            lin.bias.copy_(torch.from_numpy(params[off + din * dout:off + din * dout + dout].copy()))
        off += din * dout + dout
        layers += [lin, torch.nn.ReLU() if l < 3 else torch.nn.Sigmoid()]
    return torch.nn.Sequential(*layers)


def _flat(model, grads=False):
    out = []
    for m in model:
        if isinstance(m, torch.nn.Linear):
            out += [(m.weight.grad if grads else m.weight).detach().reshape(-1), (m.bias.grad if grads else m.bias).detach()]
    return torch.cat(out).numpy()


def _problem(orc, n, rows, seed):
    rng = np.random.default_rng(seed)
    a = orc.action_dim(n)
    dims = [2 * a, 48, 64, 48, a]
    n_params = sum(dims[l] * dims[l + 1] + dims[l + 1] for l in range(4))
    params = (rng.standard_normal(n_params) * 0.2).astype(np.float32)
    x = (rng.random((rows, 2 * a)) < 0.3).astype(np.float32)
    o = rng.random((rows, a)).astype(np.float32)
    w = (rng.random((rows, a)) < 0.15).astype(np.float32)
    return dims, params, x, o, w


def test_gradients_match_torch_autograd(orc):
    dims, params, x, o, w = _problem(orc, 9, 37, 0)
    loss, grads = orc.model_gradients(params, dims, x, o, w)
    model = _torch_model(params, dims)
    wn = torch.from_numpy(w) / torch.from_numpy(w).sum()
    tl = (((model(torch.from_numpy(x)) - torch.from_numpy(o)) ** 2) * wn).sum()
    tl.backward()
    assert abs(loss - tl.item()) <= 1e-6 * abs(tl.item())
    tg = _flat(model, grads=True)
    assert np.allclose(grads, tg, rtol=1e-4, atol=1e-7)


def test_adam_steps_match_torch_adam(orc):
    dims, params, x, o, w = _problem(orc, 9, 29, 1)
    model = _torch_model(params, dims)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6)
    st = orc.AdamState(params.size)
    cur = params.copy()
    for it in range(5):
        loss, cur = orc.update_model(cur, dims, x, o, w, st)
        opt.zero_grad()
        wn = torch.from_numpy(w) / torch.from_numpy(w).sum()
        tl = (((model(torch.from_numpy(x)) - torch.from_numpy(o)) ** 2) * wn).sum()
        tl.backward()
        opt.step()
        assert abs(loss - tl.item()) <= 1e-5 * abs(tl.item())
        assert np.allclose(cur, _flat(model), rtol=1e-5, atol=2e-7)
    assert st.t == 5
    assert np.abs(cur - params).max() > 1e-4  # the parameters did move (5 steps of lr 1e-4)
