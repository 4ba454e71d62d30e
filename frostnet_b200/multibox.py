"""SSD MultiBox loss with the reference's constructor and forward (Object_Detection/layers/modules/multibox_loss.py:9-117;
SURVEY.md 8f, row f2).  The reference matches priors to ground-truth boxes image by image in a Python loop on the CPU
(``match`` of layers/box_utils.py:71-113) and copies the targets to the device; here the whole batch is matched by ONE kernel
(csrc/multibox.cu, ``frost_multibox_match``) on the device the predictions live on.  The loss arithmetic behind it (smooth L1 on
the positives, log-sum-exp confidence loss, hard negative mining by two sorts, cross entropy) is the reference's sequence of
torch ops, unchanged - it was already device code there.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L


def match_batch(threshold, targets, priors, variances):
    """``match`` (box_utils.py:71-113) for every image of the batch: targets = list of [num_objs, 5] tensors (point-form box +
    label), priors [num_priors, 4] centre-size.  Returns loc_t [B, P, 4] float32 and conf_t [B, P] int64 on priors' device."""
    dev = priors.device
    if not priors.is_cuda:
        raise RuntimeError("frostnet_b200: MultiBox matching runs on a CUDA device (B200) only; got CPU priors")
    B, P = len(targets), priors.shape[0]
    max_obj = max(1, max(int(t.shape[0]) for t in targets))
    truths = torch.zeros((B, max_obj, 4), dtype=torch.float32, device=dev)
    labels = torch.zeros((B, max_obj), dtype=torch.int64, device=dev)
    counts = torch.tensor([int(t.shape[0]) for t in targets], dtype=torch.int32, device=dev)
    for i, t in enumerate(targets):
        if t.shape[0]:
            t = t.to(dev)
            truths[i, :t.shape[0]] = t[:, :4].float()
            labels[i, :t.shape[0]] = t[:, 4].long()
    priors = priors.detach().float().contiguous()
    loc_t = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
    conf_t = torch.empty((B, P), dtype=torch.int64, device=dev)
    ov = torch.empty((B, P), dtype=torch.float32, device=dev)
    ix = torch.empty((B, P), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.call("frost_multibox_match", truths.data_ptr(), labels.data_ptr(), counts.data_ptr(), B, max_obj, priors.data_ptr(), P,
               float(threshold), float(variances[0]), float(variances[1]), loc_t.data_ptr(), conf_t.data_ptr(), ov.data_ptr(),
               ix.data_ptr(), L.stream(dev))
    return loc_t, conf_t


class MultiBoxLoss(nn.Module):
    """multibox_loss.py:9-117.  ``variance`` replaces the reference's module-level ``from data import coco as cfg``."""

    def __init__(self, num_classes, overlap_thresh, prior_for_matching, bkg_label, neg_mining, neg_pos, neg_overlap, encode_target,
                 use_gpu=True, variance=(0.1, 0.2)):
        super().__init__()
        self.use_gpu = use_gpu
        self.num_classes = num_classes
        self.threshold = overlap_thresh
        self.background_label = bkg_label
        self.encode_target = encode_target
        self.use_prior_for_matching = prior_for_matching
        self.do_neg_mining = neg_mining
        self.negpos_ratio = neg_pos
        self.neg_overlap = neg_overlap
        self.variance = list(variance)

    def forward(self, predictions, targets):
        loc_data, conf_data, priors = predictions
        num = loc_data.size(0)
        priors = priors[:loc_data.size(1), :].to(loc_data.device)
        loc_t, conf_t = match_batch(self.threshold, targets, priors, self.variance)

        pos = conf_t > 0
        # localization loss (smooth L1) on the positives
        pos_idx = pos.unsqueeze(pos.dim()).expand_as(loc_data)
        loss_l = F.smooth_l1_loss(loc_data[pos_idx].view(-1, 4), loc_t[pos_idx].view(-1, 4), reduction="sum")
        # confidence loss of every prior for the hard negative mining
        batch_conf = conf_data.view(-1, self.num_classes)
        x_max = batch_conf.detach().max()
        lse = torch.log(torch.sum(torch.exp(batch_conf - x_max), 1, keepdim=True)) + x_max          # box_utils.py:161-172
        loss_c = lse - batch_conf.gather(1, conf_t.view(-1, 1))
        loss_c = loss_c.masked_fill(pos.view(-1, 1), 0).view(num, -1)
        _, loss_idx = loss_c.sort(1, descending=True)
        _, idx_rank = loss_idx.sort(1)
        num_pos = pos.long().sum(1, keepdim=True)
        num_neg = torch.clamp(self.negpos_ratio * num_pos, max=pos.size(1) - 1)
        neg = idx_rank < num_neg.expand_as(idx_rank)
        # confidence loss on positives + mined negatives
        sel = (pos | neg)
        conf_p = conf_data[sel.unsqueeze(2).expand_as(conf_data)].view(-1, self.num_classes)
        loss_c = F.cross_entropy(conf_p, conf_t[sel], reduction="sum")
        N = num_pos.sum().float()
        return loss_l / N, loss_c / N
