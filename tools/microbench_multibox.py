"""MultiBox target matching at SSD300 size (8732 priors, 64 images, 1-40 boxes each): ONE launch of frost_multibox_match against
the reference's structure - a Python loop over the images running `match`'s torch ops (box_utils.py:71-113) on the CPU, then the
copy of loc_t / conf_t to the device - and against the same loop with the tensors on the device.

    python tools/microbench_multibox.py
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frostnet_b200 as F  # noqa: E402


def match_torch(threshold, truths, priors, variances, labels, loc_t, conf_t, idx):
    """the reference's algorithm with plain torch ops (what its Python loop executes per image)"""
    pf = torch.cat([priors[:, :2] - priors[:, 2:] / 2, priors[:, :2] + priors[:, 2:] / 2], 1)
    mx = torch.min(truths[:, None, 2:], pf[None, :, 2:])
    mn = torch.max(truths[:, None, :2], pf[None, :, :2])
    inter = (mx - mn).clamp(min=0).prod(2)
    area_a = ((truths[:, 2] - truths[:, 0]) * (truths[:, 3] - truths[:, 1]))[:, None]
    area_b = ((pf[:, 2] - pf[:, 0]) * (pf[:, 3] - pf[:, 1]))[None, :]
    ov = inter / (area_a + area_b - inter)
    _, bpi = ov.max(1)
    bto, bti = ov.max(0)
    bto.index_fill_(0, bpi, 2)
    for j in range(bpi.shape[0]):
        bti[bpi[j]] = j
    m = truths[bti]
    conf = labels[bti].long() + 1
    conf[bto < threshold] = 0
    g_cxcy = ((m[:, :2] + m[:, 2:]) / 2 - priors[:, :2]) / (variances[0] * priors[:, 2:])
    g_wh = torch.log((m[:, 2:] - m[:, :2]) / priors[:, 2:]) / variances[1]
    loc_t[idx] = torch.cat([g_cxcy, g_wh], 1)
    conf_t[idx] = conf


def main():
    dev = "cuda:0"
    torch.manual_seed(0)
    P, B = 8732, 64
    priors = torch.cat([torch.rand(P, 2), 0.05 + torch.rand(P, 2) * 0.5], 1)
    targets = []
    for b in range(B):
        n = int(torch.randint(1, 41, (1,)))
        xy = torch.rand(n, 2) * 0.7
        targets.append(torch.cat([xy, (xy + 0.05 + torch.rand(n, 2) * 0.25).clamp(max=1.0), torch.randint(0, 20, (n, 1)).float()], 1))
    pd, td = priors.to(dev), [t.to(dev) for t in targets]

    def ours():
        return F.match_batch(0.5, td, pd, [0.1, 0.2])

    def loop(pr, tg, device):
        loc_t = torch.empty(B, P, 4, device=device)
        conf_t = torch.empty(B, P, dtype=torch.long, device=device)
        for i in range(B):
            match_torch(0.5, tg[i][:, :4], pr, [0.1, 0.2], tg[i][:, 4], loc_t, conf_t, i)
        return loc_t.to(dev), conf_t.to(dev)

    for name, fn in (("frost_multibox_match (1 launch + packing)", ours), ("per-image torch ops on the CPU + copy (the reference's structure)",
                                                                        lambda: loop(priors, targets, "cpu")),
                     ("per-image torch ops on the device", lambda: loop(pd, td, dev))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        print("%-70s %8.2f ms per batch of %d images" % (name, (time.perf_counter() - t0) / 5 * 1e3, B))


if __name__ == "__main__":
    main()
