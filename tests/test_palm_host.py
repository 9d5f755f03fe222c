"""``dots_socp_b200.palm.q_lambda_step`` (the extra q / lambda solve of ``is_palm=True``) against the oracle's
``solve_q_lambda`` on CPU tensors, through the engine's internal layout with a non-trivial vertex / triangle renumbering."""
import numpy as np
import torch

from oracle import alm_oracle as orc
from dots_socp_b200 import palm, synth


def to_vertex(x, pv):
    return torch.from_numpy(np.ascontiguousarray(x[:, pv]))


def to_tri(x, pf):
    return torch.from_numpy(np.ascontiguousarray(x[:, pf, :].transpose(0, 2, 1)))


def to_corner(x, pf):                                   # (nT, 2, 3, T, 3) -> [tau][side][k][xyz][f], Engine.to_internal
    nT, T = x.shape[0], x.shape[3]
    out = np.zeros((nT + 1, 2, 3, 3, T))
    out[:nT, 0] = x[:, 0][:, :, pf, :].transpose(0, 1, 3, 2)
    out[1:, 1] = x[:, 1][:, :, pf, :].transpose(0, 1, 3, 2)
    return torch.from_numpy(out)


def test_q_lambda_step_equals_the_oracle_on_the_internal_layout():
    geo, _ = synth.example("icosphere2")
    n_time, cong, r, s = 6, 0.07, 1.3, 2.6
    ops = orc.MeshOps(n_time, geo, build_inverse=False)
    rng = np.random.default_rng(11)
    V, T = ops.V, ops.T
    phi = rng.standard_normal((n_time + 1, V))
    mu, z_fst, z_end, b_fst, b_end = (rng.standard_normal((n_time, V)) for _ in range(5))
    E = rng.standard_normal((n_time + 1, T, 3))
    z_mid, b_mid = (rng.standard_normal((n_time, 2, 3, T, 3)) for _ in range(2))
    dt_phi, dx_phi = orc.grad_time(ops.dt, phi), orc.grad_space(ops.G, phi)
    A_ref, B_ref, lc_ref = ops.solve_q_lambda(s, cong, r, dt_phi, dx_phi, mu, E, z_fst, z_mid, z_end, b_fst, b_mid, b_end)
    pv, pf = rng.permutation(V), rng.permutation(T)     # internal numbering: any renumbering of vertices / triangles
    A, lam_c, B = palm.q_lambda_step(ops.dt, s, cong, r, to_vertex(phi, pv), to_tri(dx_phi, pf), to_vertex(mu, pv), to_tri(E, pf),
                                     to_vertex(z_fst, pv), to_vertex(z_end, pv), to_corner(z_mid, pf),
                                     to_vertex(b_fst, pv), to_vertex(b_end, pv), to_corner(b_mid, pf))
    assert np.abs(A.numpy() - A_ref[:, pv]).max() <= 1e-13 * np.abs(A_ref).max()
    assert np.abs(lam_c.numpy() - lc_ref[:, pv]).max() <= 1e-13 * max(np.abs(lc_ref).max(), 1e-300)
    assert np.abs(B.numpy() - B_ref[:, pf, :].transpose(0, 2, 1)).max() <= 1e-13 * np.abs(B_ref).max()
