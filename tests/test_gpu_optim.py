"""GradBoost kernels vs the golden vectors recorded from the reference optimizer.py (injected noise)."""
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["QSGD", "QRMS", "QRMSc", "QAdam", "QAdamW"])
def test_gradboost_matches_reference(kind):
    import frostnet_b200 as F
    g = load_golden("gradboost.pt")[kind]
    dev = torch.device("cuda:0")
    cls = {"QSGD": F.QSGD, "QRMS": F.QRMSprop, "QRMSc": F.QRMSprop, "QAdam": F.QAdam, "QAdamW": F.QAdamW}[kind]
    params = [torch.nn.Parameter(p.clone().to(dev)) for p in g["p0"]]
    opt = cls(params, clip_by=1e-3, toss_coin=True, noise_decay=1e-2, **g["kw"])
    nsteps, warm = len(g["grads"]), g["warm"]
    for t in range(nsteps):
        if t == warm:
            opt.is_warmup = False
        for p, x in zip(params, g["grads"][t]):
            p.grad = x.clone().to(dev)
        if t >= warm:
            opt.inject_noise(g["noises"][t - warm], g["coins"][t - warm])
        opt.step()
        snap = g["snaps"][t]
        for i, p in enumerate(params):
            torch.testing.assert_close(p.detach().cpu(), snap["params"][i], rtol=2e-6, atol=1e-7, msg=lambda m: "%s step %d param %d: %s" % (kind, t, i, m))
            torch.testing.assert_close(p.grad.cpu(), snap["grads_after"][i], rtol=2e-6, atol=1e-9)
            torch.testing.assert_close(opt.state[p]["exp_max"].cpu(), snap["exp_max"][i], rtol=2e-6, atol=1e-9)
            assert torch.equal(opt.state[p]["exp_min"].cpu(), snap["exp_min"][i])          # identically 0 (K9)
    for i, p in enumerate(params):
        ref = g["final_state"][i]
        st = opt.state[p]
        assert set(ref.keys()) == set(st.keys()), (kind, sorted(ref.keys()), sorted(st.keys()))
        for k, v in ref.items():
            if torch.is_tensor(v):
                torch.testing.assert_close(st[k].cpu(), v, rtol=2e-4, atol=1e-8)   # fma contraction differs between ATen CPU and nvcc
            else:
                assert st[k] == v, (k, st[k], v)


def test_device_noise_statistics():
    """Philox |Laplace| * coin: mean of the boost equals E|L|*P(coin)*sens when unclipped."""
    import frostnet_b200 as F
    dev = torch.device("cuda:0")
    n = 1 << 20
    p = torch.nn.Parameter(torch.zeros(n, device=dev))
    opt = F.QSGD([p], lr=0.0, momentum=0.0, clip_by=0.0, toss_coin=True, noise_decay=0.0)
    p.grad = torch.ones(n, device=dev)
    opt.step()                                   # warm-up: exp_max = (0.1*1)/0.1 = 1
    opt.is_warmup = False
    p.grad = torch.ones(n, device=dev)
    opt.step()
    emax = opt.state[p]["exp_max"]
    boost = (p.grad - 1.0) / emax                # |L| * coin
    coin = opt.state[p]["coin_toss"]
    assert abs(float(coin.mean()) - 0.5) < 5e-3
    nz = boost[coin > 0]
    assert abs(float(nz.mean()) - 1.0) < 1e-2     # Exp(1): mean 1, var 1
    assert abs(float(nz.var()) - 1.0) < 3e-2
    assert float(boost[coin == 0].abs().max()) == 0.0
    # KS distance against Exp(1)
    xs = torch.sort(nz).values.double().cpu()
    cdf = 1 - torch.exp(-xs)
    emp = torch.arange(1, xs.numel() + 1, dtype=torch.double) / xs.numel()
    assert float((cdf - emp).abs().max()) < 5e-3
