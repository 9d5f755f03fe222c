"""The launch list of one ALM iteration on one rank of an N-rank run, from the host-side plan alone (no GPU needed): kernel,
grid, block, and the cross-rank fences between them.  ncu cannot attach to a multi-rank job here; this is the exact sequence
the sharded engine enqueues (dots_socp_b200/engine.py: _iterate_sharded, csrc/lap_kernels.cu: launch_sweeps, csrc/sweep_ring.cu:
sr_sweeps), and the per-group durations measured live are in profiles/r2_scale_*.json.

    python tools/launch_plan.py icosphere7_nt63 8 > profiles/r2_launch_plan_icosphere7_nt63_n8.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                           # noqa: E402

from bench import WORKLOADS                                  # noqa: E402
from dots_socp_b200 import capi, dist as dd, nested, ring_plan, surface, synth      # noqa: E402
from dots_socp_b200.engine import _sweep_items               # noqa: E402

workload, world = sys.argv[1], int(sys.argv[2])
n_sm = 148
ex, n_time, cong, _ = WORKLOADS[workload]
geo, _ = synth.example(ex)
lib = capi.load()
v = np.ascontiguousarray(geo["vertices"], dtype=np.float64)
tri = np.ascontiguousarray(geo["triangles"]).astype(np.int64)
V, T = v.shape[0], tri.shape[0]
sym = nested.analyse_native(lib, v, surface.mesh_operators_native(lib, v, tri)["K"], leaf_size=16)
per_rank = -(-(n_time + 1) // world)
pad = dd.pad_modes(per_rank)
ring = not (2 * 8 * max(32, pad) * sym.panel_entries / 2 ** 20 < 300 or (world > 1 and pad % 32))
part = dd.partition(n_time, 0, world, min_pad=32 if ring else 0)
levels, steps = part.lvl_end - part.lvl_begin, part.t_end - part.lvl_begin
cdiv = lambda a, b: -(-a // b)
rows = []
add = lambda k, g, b, note="": rows.append((k, g, b, note))
add("k_phi_rhs", f"({cdiv(V, 256)}, {levels}, 1)", 256, "stores its slab into every rank's rhs (peer memory)" if world > 1 else "")
if world > 1:
    add("FENCE (1-element all-reduce)", "", "", "rhs slabs visible everywhere")
sym_tt = world == 1 and (n_time + 1) % 16 == 0 and part.m_pad == n_time + 1
add("k_time_sym<0>" if sym_tt else "k_time_mma<0>", "persistent / one tile per block", 256, f"all {n_time + 1} levels -> {part.m_pad} modes of this rank")
if ring:
    rp = ring_plan.build(sym=sym, n_sm=n_sm, m_pad=part.m_pad, split_bytes=64 * 1024, tasks_per_sm=64, task_bytes=(16 * 1024, 48 * 1024), wpr_max=8)
    for lv in range(sym.n_levels):
        gn = int(rp["gv_ptr"][lv + 1] - rp["gv_ptr"][lv])
        if gn:
            add(f"k_ring_gather<{part.m_pad}>", f"({cdiv(gn, 8)}, 1, 1)", 256, f"forward level {lv}: pull the descendants' contributions")
        n = int(rp["fwd_ptr"][lv + 1] - rp["fwd_ptr"][lv])
        w = int(rp["fwd_wpr"][lv])
        if n:
            add(f"k_ring_run<{part.m_pad}, 0>" if w == 1 else f"k_ring_split<{part.m_pad}, {w}, 0>", f"({cdiv(n, 8) if w == 1 else n}, 1, 1)", 256, f"forward level {lv}")
    for lv in range(sym.n_levels - 1, -1, -1):
        n = int(rp["bwd_ptr"][lv + 1] - rp["bwd_ptr"][lv])
        w = int(rp["bwd_wpr"][lv])
        if n:
            add(f"k_ring_run<{part.m_pad}, 1>" if w == 1 else f"k_ring_split<{part.m_pad}, {w}, 1>", f"({cdiv(n, 8) if w == 1 else n}, 1, 1)", 256, f"backward level {lv}")
else:
    plan = _sweep_items(sym, n_sm, part.m_pad)
    for lv in range(sym.n_levels):
        gn = int(plan["node_ptr"][lv + 1] - plan["node_ptr"][lv])
        fused = bool(plan["wpr"][lv] & 16)
        if lv > 0 and gn and not fused:
            add("k_sweep_gather", f"({gn}, 1, 1)", 256, f"forward level {lv}")
        n = int(plan["fwd_ptr"][lv + 1] - plan["fwd_ptr"][lv])
        if n:
            add(f"k_sweep_run<{part.m_pad}, {int(plan['wpr'][lv]) & 15}, 0, {'true' if fused else 'false'}>", f"({n}, 1, 1)", 256, f"forward level {lv}")
    for lv in range(sym.n_levels - 1, -1, -1):
        n = int(plan["bwd_ptr"][lv + 1] - plan["bwd_ptr"][lv])
        if n:
            add(f"k_sweep_run<{part.m_pad}, {int(plan['cw'][lv])}, 1, false>", f"({n}, 1, 1)", 256, f"backward level {lv}")
if world > 1:
    add("FENCE (1-element all-reduce)", "", "", "every rank's modes of the solution are final")
add("k_time_sym<1>" if sym_tt else "k_time_mma<1>", "persistent / one tile per block", 256,
    "all modes -> own levels" + (" (the other ranks' modes are loaded over NVLink)" if world > 1 else ""))
add("k_vertex", f"({cdiv(V, 256)}, {steps}, 1)", 256, "pushes the halo step to the next rank" if world > 1 else "")
if world > 1:
    add("FENCE (1-element all-reduce)", "", "", "vertex halos visible")
tiles, wave = cdiv(T, 128), 3 * n_sm
tch = 16
if tiles * cdiv(levels, 16) < 2 * wave:
    tch = 2
    while tch < 16 and tiles * cdiv(levels, tch) > wave:
        tch += 1
add("k_tri_tma<0>", f"({tiles}, {cdiv(levels, tch)}, 1)", 128, f"{tch} time levels per block" + ("; pushes the corner halo to the previous rank" if world > 1 else ""))
print("launch,kernel,grid,block,note")
for i, r in enumerate(rows):
    print(f"{i},\"{r[0]}\",\"{r[1]}\",{r[2]},\"{r[3]}\"")
print(f"# {workload}, rank 0 of {world}: {levels} time levels, {part.m_pad} padded modes, {sum(1 for r in rows if not r[0].startswith('FENCE'))} kernel launches, "
      f"{sum(1 for r in rows if r[0].startswith('FENCE'))} fences per iteration", file=sys.stderr)
