"""`validate()`-compatible evaluation loop: the CALLER of the hot path.

Mirrors `validate(val_loader, model, criterion, sparsity_criterion, args, epoch)` of the reference's
`imagenet_classification/train/main.py:607-757` - same arguments, same return tuple
`(top1, top5, loss, act_rate, flops_G, all_density[4, n_blocks])`, the same per-batch quantities (classification loss,
`act_rate = mean(flops_perc_list)`, FLOPs target loss, `flops / 1e9`, top-1 / top-5 via `accuracy`, the four density
lists weighted by batch size) - over a drop-in LAUD model whose forward runs on the CUDA path.

B200-first differences (results equal the reference's up to fp32 summation order):
  * the reference issues seven blocking `dist.all_reduce` + `.item()` round trips PER BATCH (:665-697); here every
    per-batch quantity is accumulated ON THE DEVICE as `value x batch_size` in one fp64 vector and reduced across
    ranks ONCE at the end (one NCCL all-reduce of ~10 + 4 x n_blocks numbers; equivalent because every rank weighs its
    batches by their size).  No host synchronisation inside the loop, so batches pipeline behind each other;
  * with `args.use_cuda_graph` (default on) the forward of each batch shape is captured once and replayed;
  * the reference guards the density all-reduce with `args.list_dyn_mode in ['channel', 'both']` - a list-vs-string
    compare that is always False, so its density file holds RANK-LOCAL densities (:719-730).  `reduce_density=False`
    (default) reproduces that; True averages over ranks as the code evidently intended.
"""
from __future__ import annotations

import math
import time
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """precision@k in percent, reference utils/utils.py:62-76 (same tie behaviour: torch.topk, largest, sorted)."""
    with torch.no_grad():
        maxk = max(topk)
        batch_size = target.size(0)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target.view(1, -1).expand(maxk, -1))
        return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / batch_size) for k in topk]


class SparsityCriterion_bounds(torch.nn.Module):
    """FLOPs-target regulariser evaluated by validate() (reference utils/sparsity_loss_unify.py:6-29)."""

    def __init__(self, sparsity_target, num_epochs, full_flops):
        super().__init__()
        self.sparsity_target, self.num_epochs, self.full_flops = sparsity_target, num_epochs, full_flops

    def forward(self, epoch, sparsity_list, flops):
        p = epoch / (0.33 * self.num_epochs)
        progress = math.cos(min(max(p, 0), 1) * (math.pi / 2)) ** 2
        upper = 1 - progress * (1 - self.sparsity_target)
        lower = progress * self.sparsity_target
        s = sparsity_list
        bounds = (torch.clamp(s - upper, min=0) ** 2 + torch.clamp(lower - s, min=0) ** 2).sum() / s.numel()
        return bounds + (flops / self.full_flops - self.sparsity_target) ** 2


class AverageMeter:
    """utils/utils.py:20-41"""

    def __init__(self, name, fmt=":f"):
        self.name, self.fmt = name, fmt
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(**self.__dict__)


def _arg(args, name, default):
    return getattr(args, name, default)


def validate(val_loader, model, criterion, sparsity_criterion, args, epoch, reduce_density: bool = False):
    """Returns (top1.avg, top5.avg, losses.avg, act_rates.avg, FLOPs.avg, all_density ndarray [4, n_blocks])."""
    model.eval()
    gpu = _arg(args, "gpu", None)
    if _arg(args, "device", None) is not None:           # (tests drive the host logic on CPU tensors with a stand-in model)
        dev = torch.device(args.device)
    else:
        dev = torch.device("cuda", gpu) if gpu is not None else torch.device("cuda", torch.cuda.current_device())
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    sparse = _arg(args, "sparse", True)
    lam = _arg(args, "lambda_act", 0.1)
    t_last = _arg(args, "t_last", 0.01)
    use_graph = _arg(args, "use_cuda_graph", True) and hasattr(model, "capture")
    log = _arg(args, "print_custom", None)
    graphs = {}
    acc = None            # fp64 [8 + 4*n_blocks]: sums of value x batch_size (+ the sample count)
    n_blocks = None
    t0 = time.time()
    with torch.no_grad():
        for i, (images, target) in enumerate(val_loader):
            images = images.to(dev, non_blocking=True)
            target = target.to(dev, non_blocking=True)
            bs = images.size(0)
            if use_graph:
                key = tuple(images.shape)
                g = graphs.get(key)
                if g is None:
                    g = graphs[key] = model.capture(images)
                logits, stats = g.run(images)
                r3, r2, r1, rc, perc, flops = model._engine.split_stats(stats)
                output = logits
            else:
                output, r3, r2, r1, rc, perc, flops = model(images, temperature=t_last)
            flops = flops / 1e9
            loss_cls = criterion(output.float(), target)
            if sparse:
                act_rate = perc.mean()
                loss_flops = sparsity_criterion(epoch, perc, flops)
                loss = loss_cls + lam * loss_flops
            else:
                act_rate = torch.ones((), device=dev)
                loss_flops = torch.zeros((), device=dev)
                loss = loss_cls
            acc1, acc5 = accuracy(output, target, topk=(1, 5))
            dens = torch.cat([torch.cat(list(lst)).double() for lst in (r3, r2, r1, rc)])
            if acc is None:
                n_blocks = dens.numel() // 4
                acc = torch.zeros(8 + dens.numel(), dtype=torch.float64, device=dev)
            vec = torch.stack([loss_cls.double(), loss_flops.double().reshape(()), loss.double().reshape(()),
                               act_rate.double().reshape(()), flops.double().reshape(()), acc1.double().reshape(()),
                               acc5.double().reshape(()), torch.ones((), dtype=torch.float64, device=dev)])
            acc[:8] += vec * bs
            acc[8:] += dens * bs
            if log is not None and i % 10 == 0:
                log(f"Test: [{i}/{len(val_loader)}]  ({time.time() - t0:.1f} s, no per-batch host sync)")
    if acc is None:
        raise ValueError("validate(): empty loader")
    local = acc.clone()
    if world > 1:
        dist.all_reduce(acc)              # the single collective of the evaluation
    n_all = acc[7].item()
    m = (acc[:7] / acc[7]).tolist()
    loss_cls_avg, loss_flops_avg, loss_avg, act_avg, flops_avg, top1, top5 = m
    dens_src = acc if reduce_density else local
    all_density = (dens_src[8:] / dens_src[7]).reshape(4, n_blocks).float().cpu().numpy()
    rank0 = (not (dist.is_available() and dist.is_initialized())) or dist.get_rank() == 0
    if log is not None and rank0:
        log(f" * Acc@1 {top1:.3f} Acc@5 {top5:.3f}  ({int(n_all)} samples)")
        log(f"* Conv3 spatial sparsity: {all_density[0]}")
        log(f"* Conv2 spatial sparsity: {all_density[1]}")
        log(f"* Conv1 spatial sparsity: {all_density[2]}")
        log(f"* channel sparsity: {all_density[3]}")
    return top1, top5, loss_avg, act_avg, flops_avg, all_density


def save_density(train_url: str, all_density: np.ndarray, is_best: bool) -> None:
    """The density files the reference's main loop writes after each validation (train/main.py:454-459)."""
    import os
    np.savetxt(os.path.join(train_url, "all_density_latest.txt"), all_density)
    if is_best:
        np.savetxt(os.path.join(train_url, "all_density_best.txt"), all_density)
