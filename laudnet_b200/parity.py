"""Bookkeeping for FREE-RUNNING parity checks of a gated network against a checker's trace.

Gating decisions are discrete: once ONE decision of a sample differs from the checker's (which can legitimately
happen where the checker's own keep/drop margin is within the error budget of the fp16 activations feeding the
masker), everything downstream of it in that sample is a different - equally valid - computation and is no longer
comparable.  So decisions are compared in execution order PER SAMPLE, up to and including the sample's first
differing block:

  * a differing decision in a sample that still agreed so far is explained only if the checker's margin
    |keep - drop| is at most `margin_tol` x max|logits| of that masker  -> counted in `first_flips`,
    otherwise in `unexplained` (a real failure);
  * later blocks of a diverged sample are skipped (`skipped_after_divergence`);
  * samples that never diverged (`agreeing`) must reproduce the checker's logits to tolerance.

Nothing here imports the oracle: callers pass plain tensors (tests, smoke(), bench.py's parity field).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class GateParity:
    def __init__(self, batch: int, margin_tol: float):
        self.margin_tol = margin_tol
        self.diverged = torch.zeros(batch, dtype=torch.bool)
        self.decisions = 0             # decisions compared (samples still in agreement)
        self.first_flips = 0           # differing decisions at a sample's first point of divergence, margin within budget
        self.unexplained = 0           # ... with a CLEAR margin: failures
        self.skipped_after_divergence = 0
        self.max_flip_margin = 0.0     # largest relative margin among the explained first flips
        self.max_logit_err = 0.0       # largest relative error of OUR gate logits (where recorded) on agreeing samples

    def block(self, got_mask: torch.Tensor, want_mask: torch.Tensor, want_logits: torch.Tensor,
              got_logits: Optional[torch.Tensor] = None) -> None:
        """One masker of one block.  got/want: [B, ...] 0/1 (any dtype); want_logits [B, 2G, ...] with the keep
        logits first (reference utils.py:55,122); got_logits (optional) ours, same layout."""
        B = want_mask.shape[0]
        G = want_logits.shape[1] // 2
        got = (got_mask.detach().cpu().reshape(B, -1) != 0)
        want = (want_mask.detach().cpu().reshape(B, -1) != 0)
        lg = want_logits.detach().cpu().double()
        scale = float(lg.abs().max().clamp_min(1e-30))
        margin = ((lg[:, :G] - lg[:, G:]).abs() / scale).reshape(B, -1)
        diff = got != want
        live = ~self.diverged
        self.decisions += int(live.sum()) * got.shape[1]
        self.skipped_after_divergence += int(self.diverged.sum()) * got.shape[1]
        d_live = diff & live[:, None]
        if d_live.any():
            m = margin[d_live]
            ok = m <= self.margin_tol
            self.first_flips += int(ok.sum())
            self.unexplained += int((~ok).sum())
            if ok.any():
                self.max_flip_margin = max(self.max_flip_margin, float(m[ok].max()))
        if got_logits is not None and live.any():
            ours = got_logits.detach().cpu().double()
            self.max_logit_err = max(self.max_logit_err, float((ours[live] - lg[live]).abs().max() / scale))
        self.diverged |= diff.any(dim=1)

    @property
    def agreeing(self) -> torch.Tensor:
        return ~self.diverged

    def logits_error(self, ours: torch.Tensor, want: torch.Tensor) -> float:
        """Normalised max error of the network logits over the samples whose every gate agreed (nan if none)."""
        keep = self.agreeing
        if not keep.any():
            return float("nan")
        o, w = ours.detach().double().cpu()[keep], want.detach().double().cpu()[keep]
        return float((o - w).abs().max() / w.abs().max().clamp_min(1e-30))

    def summary(self) -> dict:
        return {"decisions_compared": self.decisions, "first_flips_within_margin": self.first_flips,
                "unexplained_flips": self.unexplained, "max_flip_margin_rel": self.max_flip_margin,
                "margin_tol_rel": self.margin_tol, "samples": int(self.diverged.numel()),
                "samples_all_gates_equal": int(self.agreeing.sum()),
                "decisions_skipped_after_divergence": self.skipped_after_divergence,
                "max_gate_logit_err_rel": self.max_logit_err}


def compare_traces(keep: Sequence, traces: Sequence, batch: int, margin_tol: float) -> GateParity:
    """keep: the engine's BlockOutputs per block; traces: the checker's per-block traces (channel_mask /
    channel_logits / spatial_mask_small / spatial_logits)."""
    gp = GateParity(batch, margin_tol)
    for ko, tr in zip(keep, traces):
        if ko.channel_mask is not None:
            gp.block(ko.channel_mask, tr.channel_mask, tr.channel_logits, ko.channel_logits)
        if ko.spatial_mask_small is not None:
            gp.block(ko.spatial_mask_small, tr.spatial_mask_small, tr.spatial_logits, ko.spatial_logits)
    return gp
