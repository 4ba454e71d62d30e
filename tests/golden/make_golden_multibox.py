"""Golden vectors for the SSD MultiBox loss: the REAL reference code - layers/box_utils.py `match` and
layers/modules/multibox_loss.py `MultiBoxLoss.forward` (Object_Detection) - on the CPU for random batches: matched targets
(loc_t, conf_t), both losses and the gradients of the predictions.  The reference module imports `data.coco` for cfg['variance']
and uses a package-relative import; both are provided as stubs (variance [0.1, 0.2], Object_Detection/data/config.py:31).
Runs only in the build container; tests/golden/multibox.pt is committed.

    python tests/golden/make_golden_multibox.py
"""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = "/root/reference/Object_Detection/layers"


def load_reference():
    def load(name, path, package=None):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        if package:
            mod.__package__ = package
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    data = types.ModuleType("data")
    data.coco = {"variance": [0.1, 0.2]}
    sys.modules["data"] = data
    for pkg in ("layers", "layers.modules"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    box = load("layers.box_utils", os.path.join(ROOT, "box_utils.py"))
    loss = load("layers.modules.multibox_loss", os.path.join(ROOT, "modules", "multibox_loss.py"), package="layers.modules")
    return box, loss


def priors_grid(g):
    """SSD-style priors (centre-size, clipped to [0, 1]): three feature maps x a few aspect ratios"""
    out = []
    for f, s in ((10, 0.2), (5, 0.45), (3, 0.7)):
        for i in range(f):
            for j in range(f):
                cx, cy = (j + 0.5) / f, (i + 0.5) / f
                for ar in (1.0, 2.0, 0.5):
                    out.append([cx, cy, s * ar ** 0.5, s / ar ** 0.5])
    return torch.tensor(out).clamp_(max=1, min=0)


def main():
    box, loss = load_reference()
    g = torch.Generator().manual_seed(1882)
    priors = priors_grid(g)
    P, C = priors.shape[0], 6
    cases = []
    for ci, B in enumerate((1, 4, 7)):
        targets = []
        for b in range(B):
            n = int(torch.randint(1, 6, (1,), generator=g))
            xy = torch.rand(n, 2, generator=g) * 0.6
            wh = 0.08 + torch.rand(n, 2, generator=g) * 0.35
            lab = torch.randint(0, C - 1, (n, 1), generator=g).float()
            targets.append(torch.cat([xy, (xy + wh).clamp(max=1.0), lab], 1))
        if ci == 2:
            targets[0] = torch.cat([targets[0], targets[0][:1]], 0)       # a duplicated box: ties in both arg-maxes
        loc_t = torch.Tensor(B, P, 4)
        conf_t = torch.LongTensor(B, P)
        for i in range(B):
            box.match(0.5, targets[i][:, :-1].data, priors.data, [0.1, 0.2], targets[i][:, -1].data, loc_t, conf_t, i)
        loc_p = (torch.randn(B, P, 4, generator=g) * 0.5).requires_grad_(True)
        conf_p = (torch.randn(B, P, C, generator=g) * 2.0).requires_grad_(True)
        crit = loss.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, use_gpu=False)
        ll, lc = crit((loc_p, conf_p, priors), targets)
        (ll + lc).backward()
        cases.append(dict(targets=targets, loc_t=loc_t.clone(), conf_t=conf_t.clone(), loc_p=loc_p.detach().clone(),
                          conf_p=conf_p.detach().clone(), loss_l=float(ll), loss_c=float(lc), dloc=loc_p.grad.clone(),
                          dconf=conf_p.grad.clone()))
        print("case", ci, "B", B, "positives", int((conf_t > 0).sum()), "loss_l %.5f loss_c %.5f" % (float(ll), float(lc)))
    torch.save(dict(priors=priors, num_classes=C, cases=cases, torch=torch.__version__), os.path.join(HERE, "multibox.pt"))


if __name__ == "__main__":
    main()
