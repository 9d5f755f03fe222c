"""The extra q / lambda solve that opens every iteration when ``is_palm=True`` (reference socp/solver_socp.py:668-672:
``solve_q_lambda`` of :1044-1065 with the gradients of the PREVIOUS phi and the current z, beta, mu, E).

``is_palm`` is a solver-only knob that neither the reference's CLI nor its interface can reach (interface.py:275-284), so
this step is not part of the fused CUDA iteration: it is a handful of whole-array device operations on the engine's
internal layout between two fused iterations (about one extra pass over the corner arrays).  Device-agnostic tensor code,
which is what lets tests/test_palm_host.py check it against the oracle on the CPU."""
from __future__ import annotations

import math

import torch


def q_lambda_step(dt: float, s: float, cong: float, r: float, phi, dx_phi, mu, E, z_fst, z_end, z_mid, b_fst, b_end, b_mid):
    """Closed-form (A, lam_c, B) from the current state.  Internal layouts (include/dots_b200.h): vertex fields
    ``[t][v]``, triangle fields ``[tau][xyz][f]``, corner fields ``[tau][side][k][xyz][f]`` (tau = t + side; the two
    slots that do not exist in the reference are zero, so plain sums over ``side`` and ``k`` are the reference's
    ``decouple_adjoin_spacial``, :944-959)."""
    c1 = s * (1.0 + cong * r)
    c2 = 1.0 + 2.0 * s * c1
    memo_a = torch.diff(phi, dim=0) / dt + mu                                           # dt_phi + mu          (:1050)
    memo_b = (s / math.sqrt(3.0)) * (z_mid + b_mid).sum(dim=2).sum(dim=1)               # adjoint(z_mid+b_mid) (:1052)
    A = (1.0 / c2) * memo_a + (c1 / c2) * (z_end + b_end - z_fst - b_fst)               # (:1056-1059)
    diag_b = torch.full((phi.shape[0], 1, 1), 1.0 + 2.0 * s * s, dtype=phi.dtype, device=phi.device)   # (:194-202)
    diag_b[0] = diag_b[-1] = 1.0 + s * s
    B = (dx_phi + E + memo_b) / diag_b                                                  # (:1064)
    lam_c = (cong * r / (1.0 + cong * r)) * (memo_a - A)                                # (:1065)
    return A, lam_c, B
