"""GradBoost optimizers - the reference's ``optimizer.py`` surface on one multi-tensor sm_100a kernel.

``QSGD`` / ``QRMSprop`` / ``QAdam`` / ``QAdamW`` and ``get_optimizer`` keep the reference's constructor
signatures, ValueError checks, ``is_warmup`` attribute, per-iteration ``param_group['lr']`` protocol
and state key names (optimizer.py:6-48, :50-206, :208-359, :361-512, :514-667).  ``step()`` is ONE
kernel launch over all parameter tensors (the reference runs ~15 tiny kernels plus a host numpy
Laplace draw and an H2D copy per tensor, optimizer.py:165-204).  CUDA only: there is no CPU path.

Differences, by construction: the |Laplace(0,1)| noise and the coin toss come from a device Philox
stream keyed by (seed, tensor, element, step) instead of numpy / ``Tensor.random_`` - identical on
every data-parallel rank; ``inject_noise`` lets tests feed the reference's own draws.
"""
import ctypes as C

import torch
from torch.optim.optimizer import Optimizer, required

from . import _lib as L

__all__ = ["QSGD", "QRMSprop", "QAdam", "QAdamW", "get_optimizer"]


class _GradBoostBase(Optimizer):
    _kind = None

    def __init__(self, params, defaults):
        self.is_warmup = True
        self.seed = 0x5EED
        self.grad_scale = 1.0
        self._inject = None            # (list of noise tensors, list of coin tensors) for the next step
        self._table = None
        super().__init__(params, defaults)

    # ---- reference state layout (optimizer.py:146-152, :289-301, :437-451) ----
    def _init_state(self, p, group, state):
        z = lambda: torch.zeros_like(p.data, memory_format=torch.preserve_format)
        state['step'] = 0
        state['restart_step'] = 0
        state['exp_min'] = z()
        state['exp_max'] = z()
        if group['toss_coin']:
            state['coin_toss'] = z()
        if self._kind == "QRMS":
            state['square_avg'] = z()
            if group['momentum'] > 0:
                state['momentum_buffer'] = z()
            if group['centered']:
                state['grad_avg'] = z()
        elif self._kind in ("QAdam", "QAdamW"):
            state['exp_avg'] = z()
            state['exp_avg_sq'] = z()
            if group['amsgrad']:
                state['max_exp_avg_sq'] = z()

    def inject_noise(self, noises, coins):
        """Use these |Laplace| / coin draws (one tensor per parameter, in param_groups order) in the next step."""
        self._inject = (list(noises), list(coins))

    def _hyper(self, group):
        h = L.OptHyper()
        h.kind = L.OPT_KINDS[self._kind]
        h.is_warmup = 1 if self.is_warmup else 0
        h.toss_coin = 1 if group['toss_coin'] else 0
        h.nesterov = 1 if group.get('nesterov', False) else 0
        h.centered = 1 if group.get('centered', False) else 0
        h.amsgrad = 1 if group.get('amsgrad', False) else 0
        h.momentum = group.get('momentum', 0.0)
        h.dampening = group.get('dampening', 0.0)
        h.beta = group.get('beta', 0.9)
        b1, b2 = group.get('betas', (0.9, 0.999))
        h.beta1, h.beta2 = b1, b2
        h.eps = group.get('eps', 1e-8)
        h.alpha = group.get('alpha', 0.99)
        h.clip_by = group['clip_by']
        h.noise_decay = group['noise_decay']
        h.grad_scale = self.grad_scale
        h.seed = self.seed
        return h

    def _upload(self, key, arr, chunks, dev):
        """Descriptor table + work list to the device WITHOUT a host synchronisation: a blocking pageable copy here made
        the host wait for the whole step's kernels every iteration, so the next step's first launches (and the GPU)
        waited for Python.  Pinned staging slots, asynchronous copies; a slot is reused only after its copy has run."""
        nbytes = C.sizeof(arr)
        c = self.__dict__.setdefault("_upload_cache", {}).get(key)
        if c is None or c["nbytes"] != nbytes or c["dev"] != dev:
            c = dict(nbytes=nbytes, dev=dev, slot=0, chunks=None, ck=None,
                     pinned=[torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)],
                     tabs=[torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(4)],
                     events=[None] * 4)
            self._upload_cache[key] = c
        i = c["slot"]
        c["slot"] = (i + 1) % 4
        if c["events"][i] is not None:
            c["events"][i].synchronize()
        C.memmove(c["pinned"][i].data_ptr(), arr, nbytes)
        c["tabs"][i].copy_(c["pinned"][i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        c["events"][i] = ev
        if c["chunks"] != chunks:                  # the work list only changes when the parameter set does
            c["chunks"] = list(chunks)
            c["ck"] = torch.tensor(chunks, dtype=torch.int32).to(dev)
        return c["tabs"][i], c["ck"]

    @staticmethod
    def _hyper_key(group):
        return tuple((k, (tuple(v) if isinstance(v, (tuple, list)) else v)) for k, v in sorted(group.items())
                     if k not in ("params", "lr", "weight_decay"))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # gather (group, p) pairs; groups that share all non-(lr, wd) hyper-parameters share a launch
        batches = {}
        flat_index = 0
        for group in self.param_groups:
            key = self._hyper_key(group)
            for p in group['params']:
                idx = flat_index
                flat_index += 1
                if p.grad is None:
                    continue
                batches.setdefault(key, (group, []))[1].append((idx, p, group))
        st = None
        for key, (group0, items) in batches.items():
            dev = items[0][1].device
            L.require_cuda(items[0][1], "parameter")
            st = L.stream(dev)
            arr = (L.OptTensor * len(items))()
            chunks = []
            keep = []
            for ti, (idx, p, group) in enumerate(items):
                if p.grad.is_sparse:
                    raise RuntimeError('GradBoost optimizers do not support sparse gradients')
                state = self.state[p]
                if len(state) == 0:
                    self._init_state(p, group, state)
                state['step'] += 1
                if not self.is_warmup:
                    state['restart_step'] += 1
                g = p.grad
                if not g.is_contiguous() or not p.is_contiguous() or g.dtype != torch.float32:
                    raise RuntimeError("frostnet_b200 optimizers need contiguous fp32 parameters and gradients")
                t = arr[ti]
                t.p, t.g = p.data_ptr(), g.data_ptr()
                t.exp_min, t.exp_max = state['exp_min'].data_ptr(), state['exp_max'].data_ptr()
                t.coin_toss = state['coin_toss'].data_ptr() if group['toss_coin'] else None
                t.first_momentum = 0
                if self._kind == "QSGD":
                    if group['momentum'] != 0:
                        if 'momentum_buffer' not in state:
                            state['momentum_buffer'] = torch.empty_like(p.data)
                            t.first_momentum = 1
                        t.buf0 = state['momentum_buffer'].data_ptr()
                elif self._kind == "QRMS":
                    t.buf0 = state['square_avg'].data_ptr()
                    t.buf1 = state['momentum_buffer'].data_ptr() if group['momentum'] > 0 else None
                    t.buf2 = state['grad_avg'].data_ptr() if group['centered'] else None
                else:
                    t.buf0, t.buf1 = state['exp_avg'].data_ptr(), state['exp_avg_sq'].data_ptr()
                    t.buf2 = state['max_exp_avg_sq'].data_ptr() if group['amsgrad'] else None
                if self._inject is not None and not self.is_warmup:
                    nz = self._inject[0][idx].to(dev, torch.float32).contiguous()
                    cn = self._inject[1][idx].to(dev, torch.float32).contiguous()
                    keep += [nz, cn]
                    t.noise, t.coin = nz.data_ptr(), cn.data_ptr()
                t.n = p.numel()
                t.lr, t.weight_decay = group['lr'], group['weight_decay']
                t.step, t.restart_step = state['step'], state['restart_step']
                nchunk = (p.numel() + L.OPT_CHUNK - 1) // L.OPT_CHUNK
                chunks.extend((ti, c) for c in range(nchunk))
            tab, ck = self._upload(key, arr, chunks, dev)
            h = self._hyper(group0)
            with torch.cuda.device(dev):      # the kernel launches on the current device: make it the parameters'
                L.call("frost_gradboost_multi", tab.data_ptr(), len(items), ck.data_ptr(), len(chunks), C.byref(h), st)
            keep += [tab, ck]
            self._keepalive = keep
        if self._inject is not None and not self.is_warmup:
            self._inject = None
        return loss


class QSGD(_GradBoostBase):
    """optimizer.py:50-206."""
    _kind = "QSGD"

    def __init__(self, params, lr=required, momentum=0, dampening=0, weight_decay=0, nesterov=False, beta=0.9,
                 eps=1e-8, clip_by=1e-3, toss_coin=True, noise_decay=1e-2):
        if lr is not required and lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov,
                        beta=beta, eps=eps, clip_by=clip_by, toss_coin=toss_coin, noise_decay=noise_decay)
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, defaults)

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault('nesterov', False)


class QRMSprop(_GradBoostBase):
    """optimizer.py:208-359."""
    _kind = "QRMS"

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False, beta=0.9,
                 clip_by=1e-3, toss_coin=True, noise_decay=1e-2):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= momentum:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if not 0.0 <= alpha:
            raise ValueError("Invalid alpha value: {}".format(alpha))
        defaults = dict(lr=lr, momentum=momentum, alpha=alpha, eps=eps, centered=centered, weight_decay=weight_decay,
                        beta=beta, clip_by=clip_by, toss_coin=toss_coin, noise_decay=noise_decay)
        super().__init__(params, defaults)

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault('momentum', 0)
            group.setdefault('centered', False)


class _QAdamBase(_GradBoostBase):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, clip_by=1e-3,
                 toss_coin=True, noise_decay=1e-2):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, clip_by=clip_by,
                        toss_coin=toss_coin, noise_decay=noise_decay)
        super().__init__(params, defaults)

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault('amsgrad', False)


class QAdam(_QAdamBase):
    """optimizer.py:361-512."""
    _kind = "QAdam"


class QAdamW(_QAdamBase):
    """optimizer.py:514-667."""
    _kind = "QAdamW"

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, clip_by=1e-3,
                 toss_coin=True, noise_decay=1e-2):
        super().__init__(params, lr, betas, eps, weight_decay, amsgrad, clip_by, toss_coin, noise_decay)


def get_optimizer(optim, params_set, args):
    """optimizer.py:6-48 - same names, same ``args`` attributes."""
    if optim == "SGD":
        return torch.optim.SGD(params_set, args.learning_rate, momentum=0.9, weight_decay=args.weight_decay,
                               nesterov=args.nesterov)
    if optim == "RMS":
        return torch.optim.RMSprop(params_set, args.learning_rate, alpha=0.9, momentum=0.9, eps=1e-8,
                                   weight_decay=args.weight_decay)
    if optim == "Adam":
        return torch.optim.Adam(params_set, args.learning_rate, betas=(0.9, 0.999), eps=1e-08,
                                weight_decay=args.weight_decay)
    if optim == "AdamW":
        return torch.optim.AdamW(params_set, args.learning_rate, betas=(0.9, 0.999), eps=1e-08,
                                 weight_decay=args.weight_decay, amsgrad=args.amsgrad)
    if optim == "QSGD":
        return QSGD(params_set, args.learning_rate, momentum=0.9, weight_decay=args.weight_decay,
                    nesterov=args.nesterov, clip_by=args.clip_by, toss_coin=args.toss_coin,
                    noise_decay=args.noise_decay)
    if optim == "QRMS":
        return QRMSprop(params_set, args.learning_rate, alpha=0.9, momentum=0.9, eps=1e-8,
                        weight_decay=args.weight_decay, clip_by=args.clip_by, toss_coin=args.toss_coin,
                        noise_decay=args.noise_decay)
    if optim == "QAdam":
        return QAdam(params_set, args.learning_rate, betas=(0.9, 0.999), eps=1e-08, weight_decay=args.weight_decay,
                     amsgrad=args.amsgrad, clip_by=args.clip_by, toss_coin=args.toss_coin,
                     noise_decay=args.noise_decay)
    if optim == "QAdamW":
        return QAdamW(params_set, args.learning_rate, betas=(0.9, 0.999), eps=1e-08, weight_decay=args.weight_decay,
                      amsgrad=args.amsgrad, clip_by=args.clip_by, toss_coin=args.toss_coin,
                      noise_decay=args.noise_decay)
    # the reference falls through and raises UnboundLocalError on `return optimizer`
    raise UnboundLocalError("local variable 'optimizer' referenced before assignment")
