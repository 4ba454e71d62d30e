"""SSD MultiBox loss (frostnet_b200/multibox.py + csrc/multibox.cu) against the reference's own code run on
the CPU (Object_Detection/layers/box_utils.py `match`, layers/modules/multibox_loss.py), golden vectors from
tests/golden/make_golden_multibox.py: the matched class targets BIT-EXACT (including a duplicated ground-truth box, i.e. ties in
both arg-maxes), the encoded offsets to logf's last ulp, both losses and the gradients of the predictions to fp32 rounding."""
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_match_and_loss_against_reference(ci):
    import frostnet_b200 as F
    g = load_golden("multibox.pt")
    c = g["cases"][ci]
    priors = g["priors"].to(DEV)
    loc_t, conf_t = F.match_batch(0.5, c["targets"], priors, [0.1, 0.2])
    assert torch.equal(conf_t.cpu(), c["conf_t"])
    pos = c["conf_t"] > 0
    # the reference encodes every prior; compare where the target is used (positives) tightly and everywhere loosely
    assert torch.allclose(loc_t.cpu()[pos], c["loc_t"][pos], rtol=2e-6, atol=2e-6)
    assert torch.allclose(loc_t.cpu(), c["loc_t"], rtol=1e-5, atol=1e-5)
    crit = F.MultiBoxLoss(g["num_classes"], 0.5, True, 0, True, 3, 0.5, False)
    loc_p = c["loc_p"].to(DEV).requires_grad_(True)
    conf_p = c["conf_p"].to(DEV).requires_grad_(True)
    ll, lc = crit((loc_p, conf_p, priors), [t.to(DEV) for t in c["targets"]])
    assert abs(float(ll.detach()) - c["loss_l"]) <= 2e-6 * abs(c["loss_l"]) + 1e-6, (float(ll), c["loss_l"])
    assert abs(float(lc.detach()) - c["loss_c"]) <= 2e-6 * abs(c["loss_c"]) + 1e-6, (float(lc), c["loss_c"])
    (ll + lc).backward()
    assert torch.allclose(loc_p.grad.cpu(), c["dloc"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(conf_p.grad.cpu(), c["dconf"], rtol=1e-5, atol=1e-7)


def test_match_at_ssd300_size_and_errors():
    """8732 priors x 64 images x up to 40 boxes in one launch: equals the reference algorithm restated with torch ops on the device"""
    import frostnet_b200 as F
    torch.manual_seed(0)
    P, B = 8732, 64
    cxcy = torch.rand(P, 2, device=DEV)
    wh = 0.05 + torch.rand(P, 2, device=DEV) * 0.5
    priors = torch.cat([cxcy, wh], 1)
    targets = []
    for b in range(B):
        n = int(torch.randint(1, 41, (1,)))
        xy = torch.rand(n, 2, device=DEV) * 0.7
        targets.append(torch.cat([xy, (xy + 0.05 + torch.rand(n, 2, device=DEV) * 0.25).clamp(max=1.0),
                                  torch.randint(0, 20, (n, 1), device=DEV).float()], 1))
    loc_t, conf_t = F.match_batch(0.5, targets, priors, [0.1, 0.2])
    pf = torch.cat([priors[:, :2] - priors[:, 2:] / 2, priors[:, :2] + priors[:, 2:] / 2], 1)
    for b in (0, 17, 63):
        t = targets[b]
        a, lab = t[:, :4], t[:, 4]
        mx = torch.min(a[:, None, 2:], pf[None, :, 2:])
        mn = torch.max(a[:, None, :2], pf[None, :, :2])
        inter = (mx - mn).clamp(min=0).prod(2)
        area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None]
        area_b = ((pf[:, 2] - pf[:, 0]) * (pf[:, 3] - pf[:, 1]))[None, :]
        ov = inter / (area_a + area_b - inter)
        bpo, bpi = ov.max(1)
        bto, bti = ov.max(0)
        bto.index_fill_(0, bpi, 2)
        for j in range(bpi.shape[0]):
            bti[bpi[j]] = j
        conf = lab[bti].long() + 1
        conf[bto < 0.5] = 0
        assert torch.equal(conf_t[b], conf), b
    with pytest.raises(RuntimeError, match="CUDA device"):
        F.match_batch(0.5, [t.cpu() for t in targets], priors.cpu(), [0.1, 0.2])
