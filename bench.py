#!/usr/bin/env python
"""bench.py — k-points/s of the per-k-point Wannier-interpolation hot path (interp + diag + Kubo/MEP) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--scaling weak|strong] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over the whole Monkhorst-Pack mesh of the workload: reset, accumulate every k-point
(Fourier sums, eigensolve, gauge rotation, band sums, frequency sweeps), ONE NCCL all-reduce of the packed partial sums inside
the library when N > 1, finalize (Fermi selection + normalisation) with the response tensors read back to the host.

Headline (the one JSON line's `value`) = BASELINE.json configs[1] (C2): FeMn 3q TB model, anomalous Hall conductivity (static +
121-point hw sweep), 128^3 k-mesh; for N > 1 the mesh grows with N (128 x 128 x 128N): "scaling": "weak".  The other four BASELINE
configs are measured in the same process after the headline and reported under `configs` at their BASELINE meshes with FIXED total
work ("scaling": "strong"): C1 32^3, C3 256^3, C4 128^3 (n_hw = 200), C5 100^3 (BASELINE 200^3 = 8 M k-points takes 14 s per step on
one GPU; `--workload C5 --mesh 200` runs it).  `--no-configs` skips them.

Keys beyond the base contract: `roofline` (FP64-pipe roofline of the dominant kernel + `step_frac` for the whole step; per-kernel
times measured live with CUDA events in a serialised profiling pass), `cpu_baseline` (the C restatement of the reference algorithm
on the host cores, bounded sample), `e2e` (model uploaded from pinned host memory + tensors read back every step), `clocks`,
`stages` (k-loop / all-reduce / finalize ms, min and max over ranks), `parity_check` (N > 1: a 16^3 mesh sharded over the N ranks
against a 1-rank run), `eigensolver_yardstick` (cuSOLVER heevjBatched on the same batch shape, comparison only).
`--impl reference` times the CPU port only (the reference is Fortran+MPI: not buildable here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "k-points/s (interp+diag+Kubo/MEP)"
UNIT = "k-points/s"
BASE_MESH = {"C1": 32, "C2": 128, "C3": 256, "C4": 128, "C5": 100}       # C5: BASELINE mesh is 200 (see module docstring)


# ------------------------------------------------------------------------------------------------ workloads
def make_workload(name: str, n_gpus: int, mesh: int = 0, scaling: str = "weak"):
    """(model, params, description) — inputs in the reference's own formats, parsed by the product's host reader.
    weak: N x N x (N * gpus) (one BASELINE-size slab per GPU); strong: N^3 whatever the GPU count."""
    from mepinterp_b200 import host
    golden = os.path.join(ROOT, "tests", "golden")
    if name == "C2":      # FeMn AHC, n_hw = 121 (tests/test_FeMn_AHE.f90 parameters, mesh of BASELINE configs[1])
        model, params, _ = host.load_run_dir(os.path.join(golden, "femn_d0.900"))
        desc = "FeMn 3q TB (nw=8, nR=27) AHC static + velo/ohc sweep n_hw=121"
    elif name == "C3":    # FeMn MEP (mep_niu + kubo_mep), 256^3
        model, params, _ = host.load_run_dir(os.path.join(golden, "femn_d0.900"))
        params.do_ahc, params.do_mep, params.do_kubo, params.n_hw, params.hw_max, params.valence_bands = False, True, True, 1, 0.0, 4
        desc = "FeMn 3q TB MEP (mep_niu + kubo_mep)"
    elif name == "C4":    # FeMn Rashba optical: photo_2nd_cond + AHC over 200 frequencies
        model, params, _ = host.load_run_dir(os.path.join(golden, "femn_rashba"))
        desc = "FeMn Rashba TB photo2 + AHC n_hw=200"
    elif name == "C1":
        model, params, _ = host.load_run_dir(os.path.join(golden, "souza_phi0"))
        desc = "Souza toy TB (nw=8, nR=7) AHC + MEP"
    elif name == "C5":    # synthetic nw=64, nR=2007, AHC, T=300 K
        from mepinterp_b200 import api, models
        model = models.synthetic_tb(64, 61)
        params = api.make_params(do_ahc=True, T_kelvin=300.0)
        desc = "synthetic TB (nw=64, nR=2007) AHC"
    else:
        raise SystemExit(f"unknown workload {name}")
    n = mesh or BASE_MESH[name]
    if scaling == "weak":
        params.mp_grid = (n, n, n * n_gpus)
    else:
        params.mp_grid = (n, n, n)
    desc += ", k-mesh %dx%dx%d" % tuple(params.mp_grid)
    if name == "C5" and n != 200:
        desc += " (BASELINE mesh 200^3)"
    return model, params, desc


def hermitian_closed(model) -> bool:
    """H(-R) = H(R)^dagger for every R (mirror of hermitian_closed in mepb_api.cu): decides which pair-item format the library uses."""
    H, R = np.asarray(model.H), np.asarray(model.R).reshape(3, -1, order="F")
    tol = 4e-16 * max(float(np.abs(H).max()), 1e-300)
    for r in range(H.shape[2]):
        partner = np.where(np.all(np.abs(R + R[:, r:r + 1]) < 1e-9, axis=0))[0]
        if partner.size == 0:
            if np.abs(H[:, :, r]).max() > tol:
                return False
            continue
        if np.abs(H[:, :, r] - H[:, :, partner[0]].conj().T).max() > tol:
            return False
    return True


def flops_model(model, params):
    """Algorithmic FP64 flops per unit for each kernel class (DESIGN.md 'Kernels and rooflines'), counted on
    the work the formulation in use needs — never more than the kernels execute:

    * sparse H(R) (FeMn: 110 of 1728 entries; the reference adds the zeros, file_io.f90:732-741): Fourier terms per NON-ZERO entry;
    * general path on a mesh: the mesh-separable transform (stage A per mesh plane, B per mesh line, C per k-point with 3 n1 rows)
      instead of the nR-term sum — counting the dense 20 nR nel would credit the kernel with 40x the work it does;
    * eigensolvers (eig, tridiag, tql, zrot): matrices/s only, no roofline fraction (a flop figure for an eigensolve
      depends on the algorithm; LAPACK-convention numbers are given as context, not used for `frac`);
    * eig_rot (general path): back-transformation U = Q Z (8 nw^3) + the six zgemms of the rotation (48 nw^3) + pair sums."""
    nw, nR = model.nw, model.nR
    H = np.asarray(model.H)
    nz = (H != 0)
    nnz = int(nz.sum())
    sparse = nw in (4, 8, 16) and 2 * nnz < nw * nw * nR                  # mepb_api.cu build_sparse
    nnz_up = int(sum(nz[p, q, :].sum() for q in range(nw) for p in range(q + 1)))
    split = nw in (4, 8)                                              # thread-per-k eigensolver kernel (mepb_eig.cu)
    herm = hermitian_closed(model)                                    # mepb_api.cu hermitian_closed: 10-double items, else 20
    large = nw not in (4, 8, 16)
    # band-sum kernel: Fourier sums of the three d_aH — per entry one complex multiply (6), the H accumulate (2) and three
    # real x complex multiply-adds (12) = 20 flops (dense: per (R, element))
    # (Hermitian-closed sparse models on the nw 4 / 8 path: the kernel sums the upper triangle only and mirrors it)
    up_only = sparse and split and herm and nnz <= 256                # mepb_api.cu build_sparse / kSubMaxNnz
    interp = 20.0 * ((nnz_up if up_only else nnz) if sparse else nR * nw * nw) + (0.0 if split else 36.0 * nw ** 3)
    if params.do_gauge_trafo:
        interp += 48.0 * nw ** 3                                      # U^dagger (d_a H) U, a = 1..3
    if params.do_ahc:
        interp += 60.0 * nw * nw
    # MEP sums (mep_niu IC + LC and kubo_mep IC + LC share every product).  Per band n0 and intermediate band n:
    # (nw - 1) x [complex cross product 42 + 12 accumulate FMAs 24 + weights 4] + local term 105 + epilogue 84 = 99 per (n0, n, m) triple.  mep_sums_kernel (nw 4 / 8) skips the bands whose weights are zero — at T = 0 only max(valence_bands, occupied) bands
    # per k-point are worked on (measure() takes the occupied count from the run's own N_el); the other paths do all nw bands,
    # with the guard and the weights applied per term (130 per triple)
    mep_per_band = 0.0
    if params.do_mep or params.do_kubo:
        if split:
            mep_per_band = 99.0 * (nw - 1) ** 2
        else:
            interp += 130.0 * nw * (nw - 1) ** 2
    out = {"interp_per_k": interp, "mep_per_band": mep_per_band, "sparse": sparse, "nnz": nnz, "herm": herm,
           # context only (LAPACK-convention count of a Hermitian eigensolve with vectors); not used for a roofline fraction
           "eig_lapack_per_k": (8.0 * (nnz_up if sparse else nR * (nw * (nw + 1) // 2)) + 36.0 * nw ** 3) if split else 0.0,
           "sweep_ahc_per_item_hw": 56.0 if herm else 92.0,          # 1 complex reciprocal + 18 (36) accumulate FMAs
           # 27 complex MACs + the two-pole factor per (item, frequency); the tensor-core sweep contracts the items of the band
           # pairs (n, l) and (l, n) together (h(l,n) = -conj h(n,l)): 246 flops per PAIR of items
           "sweep_photo_per_item_hw": 123.0 if (split and params.n_hw >= 16) else 246.0,
           "fourier_per_k": 0.0, "eig_rot_per_k": 0.0, "tridiag_lapack_per_k": 0.0, "fourier_form": None}
    if large:
        nel = nw * (nw + 1) // 2 if herm else nw * nw
        R = np.asarray(model.R).reshape(3, -1, order="F")
        n1, n2, n3 = (len(np.unique(R[a])) for a in range(3))
        N1, N2, _ = (int(x) for x in params.mp_grid)
        separable = (n1 + n2 + n3 <= 96) and n2 <= 32 and n3 <= 32                         # build_sep_plan
        if separable:
            # stage A, per mesh plane: nR terms x nel, complex multiply (6) + add (2) + real x complex FMA (4) = 12
            # stage B, per mesh line:  3 n1 rows x n2 terms x nel, complex multiply (6) + real x complex FMA (4) = 10
            # stage C, per k-point:    3 n1 rows x nel, complex multiply (6) + four real x complex FMAs (16) = 22
            out["fourier_per_k"] = 12.0 * nR * nel / (N1 * N2) + 10.0 * 3 * n1 * n2 * nel / N1 + 22.0 * 3 * n1 * nel
            out["fourier_form"] = f"mesh-separable (n1,n2,n3 = {n1},{n2},{n3}): {3 * n1} rows per k-point instead of {nR}"
        else:
            out["fourier_per_k"] = 20.0 * nR * nel
            out["fourier_form"] = "dense nR-term contraction"
        out["eig_rot_per_k"] = (8.0 + (48.0 if params.do_gauge_trafo else 0.0)) * nw ** 3 + 60.0 * nw * nw
        out["tridiag_lapack_per_k"] = 16.0 / 3.0 * nw ** 3            # zhetrd convention, context only
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self) -> float:
        return time.perf_counter()

    def summary(self, t0: float = None, t1: float = None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in list(self.rows):
            if len(r) < 9 or (t0 is not None and ts < t0) or (t1 is not None and ts > t1 + 0.05):
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()


# ------------------------------------------------------------------------------------------------ CPU side
def cpu_port_rate(model, params, budget_s: float, threads: int):
    """k-points/s of the C restatement of the reference algorithm (oracle/cpu_ref.c: dense loops + zheevd + zgemm per k,
    `threads` workers with round-robin k ownership standing in for the MPI ranks) on a bounded sample of the mesh."""
    from oracle import cpu_ref
    h = cpu_ref.Handle(model, params)
    nk_mesh = int(params.mp_grid[0]) * int(params.mp_grid[1]) * int(params.mp_grid[2])
    probe = min(nk_mesh, 256 * threads)
    t0 = time.perf_counter(); h.accumulate(0, probe, nthreads=threads); t_probe = time.perf_counter() - t0
    n = int(min(nk_mesh, max(probe, probe * budget_s / max(t_probe, 1e-6))))
    # sample = a centred contiguous range of the linear mesh index (per-k cost is mesh independent)
    b0 = (nk_mesh - n) // 2
    t0 = time.perf_counter(); h.accumulate(b0, b0 + n, nthreads=threads); dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    model, params, desc = make_workload(args.workload, args.gpus, args.mesh, args.scaling or default_scaling(args.workload))
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    budget = max(2.0, min(15.0, 150.0 / max(total, 1)))
    rates, n_sample = [], 0
    for i in range(total):
        r, n_sample, dt = cpu_port_rate(model, params, budget, threads)
        if i >= args.warmup:
            rates.append((n_sample, dt))
    nk = sum(n for n, _ in rates); tt = sum(t for _, t in rates)
    v = nk / tt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tt / max(len(rates), 1), "higher_is_better": True,
            "scaling": args.scaling or default_scaling(args.workload), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "note": "CPU port of the reference algorithm (C restatement, OpenBLAS "
                       "zheevd/zgemm); the Fortran+MPI reference cannot be built in this image (no gfortran/mpifort)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n_sample} consecutive k-points of the mesh per step"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def default_scaling(workload: str) -> str:
    return "weak" if workload == "C2" else "strong"


# ------------------------------------------------------------------------------------------------ GPU arm
class Env:
    """Process-wide plumbing: rank / world, torch device, barrier, max/min over ranks."""

    def __init__(self, rank, world, local):
        import torch
        self.torch, self.rank, self.world, self.local = torch, rank, world, local
        self.dev = torch.device("cuda", local)
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)           # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op="max"):
        if self.world == 1:
            return [float(v) for v in vals]
        import torch.distributed as dist
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return [float(x) for x in t]


def measure(env: Env, workload: str, steps: int, warmup: int, scaling: str, mesh: int, sampler, cpu_budget: float,
            partition: str, want_parity: bool) -> dict:
    """One workload on all ranks: device-timed value, e2e, clocks of the timed region, per-kernel profile, roofline."""
    import mepinterp_b200 as mb
    from mepinterp_b200 import dist as mdist
    torch, rank, world = env.torch, env.rank, env.world
    model, params, desc = make_workload(workload, world, mesh, scaling)
    nk_total = int(params.mp_grid[0]) * int(params.mp_grid[1]) * int(params.mp_grid[2])
    ctx = mb.Context(model, params, device=env.local)
    mdist.init_comm(ctx, rank, world)
    H_pinned = torch.empty(model.H.shape[::-1], dtype=torch.complex128).pin_memory()      # (nR, nw, nw) C-order == Fortran (nw,nw,nR)
    H_host = H_pinned.numpy().T                                              # Fortran-ordered view on pinned memory
    H_host[...] = model.H

    def step():
        return mdist.run_mesh_sharded(ctx, rank, world, partition=partition)

    def step_e2e():
        ctx.update_model(H_host)                                             # H2D from pinned host memory, every step
        return mdist.run_mesh_sharded(ctx, rank, world, partition=partition)  # finalize reads the tensors back (D2H)

    parity = None
    if want_parity and world > 1:
        # correctness of the N-rank path, carried in the bench line: a 16^3 mesh sharded over the N ranks (round-robin blocks,
        # library all-reduce) against the same context run on one rank
        import copy
        p16 = copy.copy(params); p16.mp_grid = (16, 16, 16)
        os.environ["MEPB_CHUNK_LOG2"] = "10"
        c16 = mb.Context(model, p16, device=env.local)
        os.environ.pop("MEPB_CHUNK_LOG2", None)
        mdist.init_comm(c16, rank, world)
        sharded = mdist.run_mesh_sharded(c16, rank, world, partition=partition)
        single = c16.run_mesh()
        c16.close()
        worst = 0.0
        for k, v in single.items():
            if isinstance(v, np.ndarray) and v.size:
                worst = max(worst, float(np.abs(np.asarray(sharded[k]) - v).max()) / max(float(np.abs(v).max()), 1e-300))
        worst = env.reduce([worst], "max")[0]
        parity = {"mesh": "16x16x16 sharded over %d ranks vs 1 rank" % world, "max_rel": worst, "n_k_ok": bool(sharded["n_k"] == 4096)}

    for _ in range(max(warmup, 3)):
        res = step()
    launches0 = ctx.kernel_launches
    t_dev, stage_acc = 0.0, {"accumulate_ms": 0.0, "allreduce_ms": 0.0, "finalize_ms": 0.0}
    t_mark0 = sampler.mark() if sampler else None
    for _ in range(steps):
        env.flush.zero_()                                                    # L2 flush between timed iterations (untimed)
        env.barrier()
        ctx.timer_start()
        res = step()
        t_dev += ctx.timer_stop()
        for k, v in ctx.timings().items():
            stage_acc[k] += v
        env.barrier()
    launches = ctx.kernel_launches - launches0
    t_e2e = 0.0
    for _ in range(steps):
        env.flush.zero_()
        env.barrier()
        ctx.timer_start()
        res = step_e2e()
        t_e2e += ctx.timer_stop()
        env.barrier()
    t_mark1 = sampler.mark() if sampler else None
    t_dev, t_e2e = env.reduce([t_dev, t_e2e], "max")
    launches = int(env.reduce([launches], "sum")[0])
    keys = list(stage_acc)
    smax = env.reduce([stage_acc[k] / steps for k in keys], "max")
    smin = env.reduce([stage_acc[k] / steps for k in keys], "min")
    stages = {k: {"min": lo, "max": hi} for k, lo, hi in zip(keys, smin, smax)}
    assert res["n_k"] == nk_total, (res["n_k"], nk_total)

    # ---- roofline: serialised profiling pass on this rank's share (never used for `value`)
    block = ctx.balanced_block(nk_total, world)

    def my_share():
        ctx.reset()
        if partition == "blocks":
            ctx.accumulate_strided(0, nk_total, block, rank, world)
        else:
            ctx.accumulate(*mdist.slab(rank, world, nk_total))

    for _ in range(2):
        ctx.set_profiling(True)
        my_share()
    prof = ctx.get_profile()
    part = ctx.get_partials()
    nk_loc = int(part[0])
    items = ctx.work_items()
    ctx.set_profiling(False)
    info = ctx.model_info()
    fm = flops_model(model, params)
    if fm["mep_per_band"]:
        # bands mep_sums_kernel works on per k-point, a lower bound: the union of {n0 <= valence_bands} (mep_niu) and, for kubo_mep,
        # the bands with f(E) != 0 — all of them at T > 0, the occupied ones at T = 0 (N_el of this rank's share / its k-points)
        zero_T = not (params.T_kelvin > 1e-2)
        occ = (float(part[2]) / max(nk_loc, 1)) if zero_T else float(model.nw)
        bands = max(float(params.valence_bands) if params.do_mep else 0.0, occ if params.do_kubo else 0.0)
        fm["interp_per_k"] += fm["mep_per_band"] * min(bands, float(model.nw))
        fm["mep_bands_per_k"] = bands
    work = {"interp": fm["interp_per_k"] * nk_loc, "fourier": fm["fourier_per_k"] * nk_loc, "eig_rot": fm["eig_rot_per_k"] * nk_loc,
            "sweep_ahc": fm["sweep_ahc_per_item_hw"] * items["sweep_ahc"] * params.n_hw,
            "sweep_photo": fm["sweep_photo_per_item_hw"] * items["sweep_photo"] * params.n_hw}
    eigen = ("eig", "tridiag", "tql", "zrot")
    kernels = {}
    for k in ("fourier", "tridiag", "tql", "zrot", "eig_rot", "eig", "interp", "sweep_ahc", "sweep_photo", "reduce"):
        if prof[k]["launches"]:
            ms = prof[k]["ms"]
            ent = {"ms_per_pass": ms, "launches": prof[k]["launches"], "share": None,
                   "tflops": (work[k] / (ms * 1e-3) / 1e12) if (k in work and work[k] > 0) else None}
            if k in eigen:
                ent["matrices_per_s"] = nk_loc / (ms * 1e-3)
            kernels[k] = ent
    tot_ms = sum(v["ms_per_pass"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = v["ms_per_pass"] / tot_ms
    rated = [k for k in kernels if kernels[k]["tflops"] is not None]
    top = max(rated, key=lambda k: kernels[k]["ms_per_pass"])
    biggest = max(kernels, key=lambda k: kernels[k]["ms_per_pass"])
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS_FP64.json")))
    peak = float(peaks["fp64_peak_tflops"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(workload, {}).get(top)
    # the frequency sweeps (n_hw >= 16) and the general path's rotation (nw % 8 == 0) run on the FP64 tensor cores (DMMA), which
    # share the FP64 pipe and its 37.1 TF peak with DFMA (tools/fp64_mixed.cu); the other kernels are plain FP64 FMA code
    dmma = (top in ("sweep_ahc", "sweep_photo") and params.n_hw >= 16) or (top == "eig_rot" and model.nw % 8 == 0)
    step_ms = t_dev / steps
    flops_step = sum(work[k] for k in work if k in kernels)
    roofline = {"kernel": top, "bound": "tensor" if dmma else "fp64", "pipe": "FP64 (DFMA and DMMA share one pipe on B200)",
                "achieved": kernels[top]["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": kernels[top]["tflops"] / peak, "traffic": traffic,
                # whole step: algorithmic flops of every kernel with a flop model (this rank's share) / its device time / peak
                "step_frac": flops_step / (tot_ms * 1e-3) / 1e12 / peak,
                "step_algorithmic_gflop": flops_step / 1e9,
                "traffic_note": "DRAM bytes per launch of the dominant kernel (ncu --set full, profiles/traffic.json); the path is "
                                "FP64-pipe bound, HBM < 10 % of its peak in every kernel",
                "peak_source": "MEASURED_PEAKS_FP64.json (tools/fp64_peaks.cu on this pool's B200: DFMA = DMMA = 37.1 TF; "
                               "MEASURED_PEAKS.json has no FP64 figure; HBM is not a bound for this path)",
                "avg_launch_ms": kernels[top]["ms_per_pass"] / kernels[top]["launches"],
                "algorithmic_flops_per_launch": work[top] / kernels[top]["launches"],
                "fourier_form": fm["fourier_form"],
                "note": None if biggest == top else f"largest share of the step is the eigensolver kernel '{biggest}' "
                                                    f"({kernels[biggest]['share']:.0%}): rated in matrices/s, no flop roofline (SURVEY 8(d))",
                "kernels": kernels}
    out = {"workload": f"{workload}: {desc}", "scaling": scaling, "k_points_per_step": nk_total, "steps": steps,
           "value": nk_total * steps / (t_dev * 1e-3), "ms_per_step": step_ms,
           "e2e": {"value": nk_total * steps / (t_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(model.H.nbytes) * world,
                   "d2h_bytes_per_step": int(ctx.partials_size() * 8 + 16) * world, "ms_per_step": t_e2e / steps},
           "gpu_launches": launches, "stages": stages, "partition": partition, "model_info": info,
           "clocks": sampler.summary(t_mark0, t_mark1) if (sampler and rank == 0) else None, "roofline": roofline}
    if parity is not None:
        out["parity_check"] = parity
    if world == 1 and cpu_budget > 0 and rank == 0:
        threads = os.cpu_count() or 1
        r, n, dt = cpu_port_rate(model, params, cpu_budget, threads)
        out["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{n} consecutive k-points of the same mesh ({dt:.1f} s), C restatement of the reference "
                                         "algorithm (dense loops + OpenBLAS zheevd/zgemm per k), one worker per core"}
    ctx.close()
    return out


def eigensolver_yardstick(nw: int, batch: int):
    """cuSOLVER heevjBatched on the batch shape of the workload (tools/heevj_compare, comparison only: libmepb never links cuSOLVER).
    A batch cuSOLVER refuses is halved (down to batch / 64) and the batch actually timed is reported."""
    exe = os.path.join(ROOT, "tools", "heevj_compare")
    if not os.path.exists(exe):
        return {"unavailable": "tools/heevj_compare not built"}
    last = ""
    b = batch
    while b >= max(1, batch // 64):
        try:
            r = subprocess.run([exe, str(nw), str(b), "2"], capture_output=True, text=True, timeout=240)
            lines = r.stdout.strip().splitlines()
            if r.returncode == 0 and lines:
                out = json.loads(lines[-1])
                if b != batch:
                    out["note"] = f"batch {batch} refused ({last.strip()[:120]}); timed on {b}"
                return out
            last = r.stderr or f"exit {r.returncode}"
        except Exception as e:                               # noqa: BLE001 — the yardstick must never break the bench line
            last = f"{type(e).__name__}: {e}"
        b //= 2
    return {"unavailable": last.strip()[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--mesh", type=int, default=0, help="override the k-mesh: N^3 (strong) or N x N x (N * gpus) (weak)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="default: weak for C2, strong for the others")
    ap.add_argument("--partition", default="blocks", choices=["blocks", "slabs"], help="k ownership: round-robin blocks or slabs")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only (skip the other BASELINE configs)")
    ap.add_argument("--no-yardstick", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
        args.gpus = world
    args.warmup = max(args.warmup, 3)
    scaling = args.scaling or default_scaling(args.workload)

    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner at communicator creation)
    # are sent to stderr for the duration of the run; the saved descriptor is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    env = Env(rank, world, local)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    head = measure(env, args.workload, args.steps, args.warmup, scaling, args.mesh, sampler,
                   0.0 if args.no_cpu_baseline else 15.0, args.partition, want_parity=True)
    extras = []
    if not args.no_configs and args.workload == "C2" and not args.mesh:
        for name in ("C1", "C3", "C4", "C5"):
            # enough steps for >= ~1 s of timed region on the small configs (clock samples need it); 2 for the long C5 step
            steps = {"C1": 400, "C3": 8, "C4": 8, "C5": 2}[name]
            try:
                extras.append(measure(env, name, steps, 3, "strong", 0, sampler, 0.0 if args.no_cpu_baseline else 4.0, args.partition,
                                      want_parity=False))
            except Exception as e:                           # noqa: BLE001 — an extra config must not lose the headline
                extras.append({"workload": name, "error": f"{type(e).__name__}: {e}"})
    yard = None
    if rank == 0 and world == 1 and not args.no_yardstick:
        yard = {}
        k = head["roofline"]["kernels"]
        if "eig" in k:
            yard["nw8"] = {"ours_eig_small_hq": {"matrices_per_s": k["eig"]["matrices_per_s"],
                                                 "note": "thread-per-k Householder + QL, includes the Fourier sum of H(k)"},
                           "cusolver": eigensolver_yardstick(8, 1 << 21)}
        c5 = next((e for e in extras if str(e.get("workload", "")).startswith("C5") and "roofline" in e), None)
        if c5:
            kk = c5["roofline"]["kernels"]
            ms = sum(kk[x]["ms_per_pass"] for x in ("tridiag", "tql", "zrot") if x in kk)
            yard["nw64"] = {"ours_tridiag_tql_zrot": {"matrices_per_s": (c5["k_points_per_step"] / (ms * 1e-3)) if ms else None,
                                                      "note": "eigenpairs of the tridiagonal form; the back-transformation U = QZ is in eig_rot"},
                            "cusolver": eigensolver_yardstick(64, 8192)}
    if sampler:
        sampler.stop()

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": head["workload"], "k_points_per_step": head["k_points_per_step"],
                           "parallelism": f"k-mesh round-robin {args.partition} x{world}, one NCCL all-reduce of the packed partial sums inside libmepb",
                           "l2": "512 MiB device buffer zeroed between timed steps (L2 flush); per-step intermediates (pair items) exceed L2",
                           "timing": "CUDA events on the library's launch stream per step, summed over steps, max over ranks"},
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "roofline": head["roofline"],
                "stages": head["stages"], "model_info": head["model_info"]}
        for k in ("parity_check", "cpu_baseline"):
            if k in head:
                line[k] = head[k]
        if extras:
            line["configs"] = extras
        if yard:
            line["eigensolver_yardstick"] = yard
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
