#!/usr/bin/env python
"""bench.py — k-points/s of the per-k-point Wannier-interpolation hot path (interp + diag + Kubo/MEP) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over the whole Monkhorst-Pack mesh of the workload: reset, accumulate every k-point
(Fourier sums, eigensolve, gauge rotation, band sums, frequency sweeps), one all-reduce of the packed partial sums when
N > 1, finalize (Fermi selection + normalisation) with the response tensors read back to the host.  Default workload =
BASELINE.json configs[1]: FeMn 3q TB model, anomalous Hall conductivity (static + 121-point hw sweep), 128^3 k-mesh.
For N > 1 the mesh grows with N (128 x 128 x 128N, each rank owns one 128^3 slab): "scaling": "weak".

Keys beyond the base contract: `roofline` (FP64-pipe roofline of the dominant kernel, per-kernel times measured live with
CUDA events in a serialised profiling pass), `cpu_baseline` (the C restatement of the reference algorithm on the host
cores, bounded sample), `e2e` (model uploaded from pinned host memory + tensors read back every step), `clocks`.
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


# ------------------------------------------------------------------------------------------------ workloads
def make_workload(name: str, n_gpus: int, mesh: int = 0):
    """(model, params, description) — inputs in the reference's own formats, parsed by the product's host reader."""
    from mepinterp_b200 import host
    golden = os.path.join(ROOT, "tests", "golden")
    if name == "C2":      # FeMn AHC, 128^3, n_hw = 121 (tests/test_FeMn_AHE.f90 parameters, mesh of BASELINE configs[1])
        model, params, _ = host.load_run_dir(os.path.join(golden, "femn_d0.900"))
        params.mp_grid = (128, 128, 128 * n_gpus)
        desc = "FeMn 3q TB (nw=8, nR=27) AHC static + velo/ohc sweep n_hw=121, k-mesh 128x128x%d" % (128 * n_gpus)
    elif name == "C3":    # FeMn MEP (mep_niu + kubo_mep), 256^3 on 8 GPUs -> 256x256x(32 N) keeps 2^21 k-points per GPU
        model, params, _ = host.load_run_dir(os.path.join(golden, "femn_d0.900"))
        params.do_ahc, params.do_mep, params.do_kubo, params.n_hw, params.hw_max, params.valence_bands = False, True, True, 1, 0.0, 4
        params.mp_grid = (256, 256, 32 * n_gpus)
        desc = "FeMn 3q TB MEP (mep_niu + kubo_mep), k-mesh 256x256x%d" % (32 * n_gpus)
    elif name == "C4":    # FeMn Rashba optical: photo_2nd_cond + AHC over 200 frequencies, 128^3
        model, params, _ = host.load_run_dir(os.path.join(golden, "femn_rashba"))
        params.mp_grid = (128, 128, 128 * n_gpus)
        desc = "FeMn Rashba TB photo2 + AHC n_hw=200, k-mesh 128x128x%d" % (128 * n_gpus)
    elif name == "C1":
        model, params, _ = host.load_run_dir(os.path.join(golden, "souza_phi0"))
        params.mp_grid = (32, 32, 32 * n_gpus)
        desc = "Souza toy TB (nw=8, nR=7) AHC + MEP, k-mesh 32x32x%d" % (32 * n_gpus)
    elif name == "C5":    # synthetic nw=64, nR=2007, AHC, T=300 K; BASELINE mesh is 200^3 (8e6 k-points, ~minutes): default 32^3 here
        from mepinterp_b200 import api, models
        model = models.synthetic_tb(64, 61)
        params = api.make_params(mp_grid=(32, 32, 32 * n_gpus), do_ahc=True, T_kelvin=300.0)
        desc = "synthetic TB (nw=64, nR=2007) AHC, k-mesh 32x32x%d (BASELINE mesh 200^3)" % (32 * n_gpus)
    else:
        raise SystemExit(f"unknown workload {name}")
    if mesh:
        params.mp_grid = (mesh, mesh, mesh * n_gpus)
        desc = desc.split(", k-mesh")[0] + ", k-mesh %dx%dx%d" % params.mp_grid
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
    """Algorithmic FP64 flops per unit for each kernel class (DESIGN.md 'Kernels and rooflines').

    Sparse H(R) (FeMn: 110 of 1728 entries): the library sums only the non-zero entries (the reference adds the zeros,
    file_io.f90:732-741 keeps them in H_tb), so the Fourier terms are counted per NON-ZERO entry — counting the dense
    8*nR*nw^2 of SURVEY 8(d) would credit the kernels with work they do not do."""
    nw, nR = model.nw, model.nR
    H = np.asarray(model.H)
    nz = (H != 0)
    nnz = int(nz.sum())
    sparse = nw in (4, 8, 16) and 2 * nnz < nw * nw * nR                  # mepb_api.cu build_sparse
    nnz_up = int(sum(nz[p, q, :].sum() for q in range(nw) for p in range(q + 1)))
    split = nw in (4, 8)                                              # thread-per-k eigensolver kernel (mepb_eig.cu)
    # eigensolver kernel: Fourier sum of H (upper triangle; 1 complex MAC = 8 flops) + LAPACK-convention 36 nw^3 eigensolve
    eig = 8.0 * (nnz_up if sparse else nR * (nw * (nw + 1) // 2)) + 36.0 * nw ** 3
    # band-sum kernel: Fourier sums of the three d_aH — per entry one complex multiply (6), the H accumulate (2) and three
    # real x complex multiply-adds (12) = 20 flops (dense: per (R, element))
    interp = 20.0 * (nnz if sparse else nR * nw * nw) + (0.0 if split else 36.0 * nw ** 3)
    if params.do_gauge_trafo:
        interp += 48.0 * nw ** 3                                      # U^dagger (d_a H) U, a = 1..3
    if params.do_ahc:
        interp += 60.0 * nw * nw
    if params.do_mep or params.do_kubo:
        interp += 130.0 * nw * (nw - 1) ** 2
    herm = hermitian_closed(model)                                    # mepb_api.cu hermitian_closed: 10-double items, else 20
    large = nw not in (4, 8, 16)
    nel = nw * (nw + 1) // 2
    return {"interp_per_k": interp, "eig_per_k": eig if split else 0.0,
            # dense Fourier contraction, Hermitian-closed input: per (k, upper-triangle element, R) one complex multiply (6),
            # one complex add (2) and three real x complex multiply-adds (12) = 20 flops (SURVEY 8(d)'s 8*nR*nw^2*4 counts the
            # full matrix with four independent complex MACs: 32 per (k, element, R) on nw^2 elements)
            "fourier_per_k": 20.0 * nR * nel if large else 0.0,
            "eig_rot_per_k": (36.0 + 48.0) * nw ** 3 + 60.0 * nw * nw if large else 0.0,
            "sweep_ahc_per_item_hw": 56.0 if herm else 92.0,          # 1 complex reciprocal + 18 (36) accumulate FMAs
            # 27 complex MACs + the two-pole factor per (item, frequency); the tensor-core sweep contracts the items of the band
            # pairs (n, l) and (l, n) together (h(l,n) = -conj h(n,l)): 246 flops per PAIR of items
            "sweep_photo_per_item_hw": 123.0 if (split and params.n_hw >= 16) else 246.0,
            "sparse": sparse, "nnz": nnz, "herm": herm}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
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
    model, params, desc = make_workload(args.workload, args.gpus, args.mesh)
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
            "warmup": args.warmup, "ms_per_step": 1e3 * tt / max(len(rates), 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "note": "CPU port of the reference algorithm (C restatement, OpenBLAS "
                       "zheevd/zgemm); the Fortran+MPI reference cannot be built in this image (no gfortran/mpifort)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n_sample} consecutive k-points of the mesh per step"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--mesh", type=int, default=0, help="override the k-mesh: N x N x (N * gpus)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner at communicator creation)
    # are sent to stderr for the duration of the run; the saved descriptor is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import mepinterp_b200 as mb
    from mepinterp_b200 import dist as mdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model, params, desc = make_workload(args.workload, world, args.mesh)
    nk_total = int(params.mp_grid[0]) * int(params.mp_grid[1]) * int(params.mp_grid[2])
    ctx = mb.Context(model, params, device=local)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)           # > 126 MB L2
    H_pinned = torch.empty(model.H.shape[::-1], dtype=torch.complex128).pin_memory()      # (nR, nw, nw) C-order == Fortran (nw,nw,nR)
    H_host = H_pinned.numpy().T                                              # Fortran-ordered view on pinned memory
    H_host[...] = model.H

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return mdist.run_mesh_sharded(ctx, rank, world)

    def step_e2e():
        ctx.update_model(H_host)                                             # H2D from pinned host memory, every step
        return mdist.run_mesh_sharded(ctx, rank, world)                      # finalize reads the tensors back (D2H)

    for _ in range(args.warmup):
        res = step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    t_dev = 0.0
    for _ in range(args.steps):
        flush.zero_()                                                        # L2 flush between timed iterations (untimed)
        barrier()
        ctx.timer_start()
        res = step()
        t_dev += ctx.timer_stop()
        barrier()
    launches = ctx.kernel_launches - launches0
    t_e2e = 0.0
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        ctx.timer_start()
        res = step_e2e()
        t_e2e += ctx.timer_stop()
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt[0])
    assert res["n_k"] == nk_total, (res["n_k"], nk_total)

    # ---- roofline of the dominant kernel: serialised profiling pass on this rank's slab (never used for `value`)
    b0, b1 = mdist.slab(rank, world, nk_total)
    ctx.set_profiling(True)
    for _ in range(2):
        ctx.set_profiling(True)
        ctx.reset(); ctx.accumulate(b0, b1)
    prof = ctx.get_profile()
    ctx.get_partials()
    items = ctx.work_items()
    ctx.set_profiling(False)
    fm = flops_model(model, params)
    nk_loc = b1 - b0
    work = {"interp": fm["interp_per_k"] * nk_loc, "eig": fm["eig_per_k"] * nk_loc,
            "fourier": fm["fourier_per_k"] * nk_loc, "eig_rot": fm["eig_rot_per_k"] * nk_loc,
            "sweep_ahc": fm["sweep_ahc_per_item_hw"] * items["sweep_ahc"] * params.n_hw,
            "sweep_photo": fm["sweep_photo_per_item_hw"] * items["sweep_photo"] * params.n_hw}
    kernels = {}
    for k in ("fourier", "tridiag", "tql", "zrot", "eig_rot", "eig", "interp", "sweep_ahc", "sweep_photo", "reduce"):
        if prof[k]["launches"]:
            ms = prof[k]["ms"]
            kernels[k] = {"ms_per_pass": ms, "launches": prof[k]["launches"], "share": None,
                          "tflops": (work[k] / (ms * 1e-3) / 1e12) if k in work else None}
    tot_ms = sum(v["ms_per_pass"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = v["ms_per_pass"] / tot_ms
    top = max((k for k in kernels if k in work), key=lambda k: kernels[k]["ms_per_pass"])
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS_FP64.json")))
    peak = float(peaks["fp64_peak_tflops"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get(top)
    # the frequency sweeps (n_hw >= 16) and the general path's rotation (nw % 8 == 0) run on the FP64 tensor cores (DMMA), which
    # share the FP64 pipe and its 37.1 TF peak with DFMA (tools/fp64_mixed.cu); the other kernels are plain FP64 FMA code
    dmma = (top in ("sweep_ahc", "sweep_photo") and params.n_hw >= 16) or (top == "eig_rot" and model.nw % 8 == 0)
    roofline = {"kernel": top, "bound": "tensor" if dmma else "fp64", "pipe": "FP64 (DFMA and DMMA share one pipe on B200)",
                "achieved": kernels[top]["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": kernels[top]["tflops"] / peak, "traffic": traffic,
                "traffic_note": "DRAM bytes per launch of the dominant kernel (ncu --set full, profiles/traffic.json); the path is "
                                "FP64-pipe bound, HBM < 10 % of its peak in every kernel",
                "peak_source": "MEASURED_PEAKS_FP64.json (tools/fp64_peaks.cu on this pool's B200: DFMA = DMMA = 37.1 TF; "
                               "MEASURED_PEAKS.json has no FP64 figure; HBM is not a bound for this path)",
                "avg_launch_ms": kernels[top]["ms_per_pass"] / kernels[top]["launches"],
                "algorithmic_flops_per_launch": work[top] / kernels[top]["launches"],
                "kernels": kernels}

    if rank == 0:
        value = nk_total * args.steps / (t_dev * 1e-3)
        e2e_v = nk_total * args.steps / (t_e2e * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {desc}", "k_points_per_step": nk_total, "parallelism": f"k-mesh slabs x{world}",
                           "l2": "512 MiB device buffer zeroed between timed steps (L2 flush); per-step intermediates (pair items) exceed L2",
                           "timing": "CUDA events on the library's launch stream per step, summed over steps, max over ranks"},
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(model.H.nbytes) * world,
                        "d2h_bytes_per_step": int(ctx.partials_size() * 8 + 16) * world, "ms_per_step": t_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r, n, dt = cpu_port_rate(model, params, 15.0, threads)
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n} consecutive k-points of the same mesh ({dt:.1f} s), C restatement of the reference "
                                              "algorithm (dense loops + OpenBLAS zheevd/zgemm per k), one worker per core"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
