"""Developer check (GPU box): libmepb vs the C oracle on a few cases; prints relative errors."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import mep_oracle as mo, cpu_ref
from tests.cases import load_case, synthetic_model, ev, rel_err
import mepinterp_b200 as mb


def compare(model, params, label, threads=8):
    h = cpu_ref.Handle(model, params)
    t = time.time(); ref = h.run_mesh(nthreads=threads); t_cpu = time.time() - t
    ctx = mb.Context(model, params)
    t = time.time(); out = ctx.run_mesh(); t_gpu = time.time() - t
    nk = ctx.n_kpts
    print(f"== {label}: nk={nk} cpu {t_cpu:.2f}s ({nk/t_cpu:.0f} k/s, {threads} thr) gpu wall {t_gpu:.3f}s kernel {ctx.last_kernel_ms:.2f} ms ({nk/ctx.last_kernel_ms*1e3:.3e} k/s)")
    for k in ref:
        if isinstance(ref[k], np.ndarray):
            print(f"   {k:14s} rel {rel_err(out[k], ref[k]):.2e}  max|ref| {np.abs(ref[k]).max():.3e}")
        else:
            print(f"   {k:14s} gpu {out[k]} ref {ref[k]}")
    ctx.close()


def main():
    rng = np.random.default_rng(0)
    # Fourier + eigen parity on an explicit list
    for name in ("femn_d0.900", "souza_phi0", "femn_rashba"):
        model, params = load_case(name)
        kp = rng.uniform(-0.5, 0.5, size=(257, 3))
        h = cpu_ref.Handle(model, params)
        e_r, V_r, Hk_r, Hka_r = h.interp(kp, want_V=True, want_H=True)
        ctx = mb.Context(model, params)
        Hk, Hka = ctx.fourier(kp)
        e, V = ctx.interp(kp)
        print(f"-- {name}: Hk {rel_err(Hk, Hk_r):.2e} Hka {rel_err(Hka, Hka_r):.2e} e {rel_err(e, e_r):.2e}")
        # gauge-invariant: |V| is invariant for non-degenerate bands ; sum_nm V V^dagger too
        print(f"   |V| {rel_err(np.abs(V), np.abs(V_r)):.2e}  diagV {rel_err(np.einsum('annk->ank', V).real, np.einsum('annk->ank', V_r).real):.2e}")
        ctx.close()
    m, p = load_case("femn_d0.900", mp_grid=(16, 16, 16))
    compare(m, p, "femn ahc 16^3")
    m, p = load_case("souza_phi0", mp_grid=(16, 16, 16), do_kubo=True)
    compare(m, p, "souza ahc+mep+kubo 16^3")
    m, p = load_case("femn_rashba", mp_grid=(8, 8, 8), n_hw=20)
    compare(m, p, "rashba opt+ahc 8^3")
    m, p = load_case("femn_d0.900", mp_grid=(12, 12, 12), do_mep=True, do_kubo=True, do_ahc=True, do_opt=True, n_hw=7, hw_min=ev(0.03), hw_max=ev(6.0),
                     n_eF=3, eF_min=ev(-0.5), eF_max=ev(0.5), T_kelvin=300.0)
    compare(m, p, "femn all jobs T=300 n_eF=3 12^3")
    sm = synthetic_model(8, 5, tb_centers=True)
    p = mo.Params(mp_grid=(10, 11, 12), do_ahc=True, do_mep=True, do_kubo=True, do_opt=True, n_hw=9, hw_min=ev(0.05), hw_max=ev(4.0), eta=ev(0.1),
                  valence_bands=3, T_kelvin=0.0, eF_min=0.0, eF_max=0.0, use_cart_velo=True)
    compare(sm, p, "synthetic nw=8 tb cart 10x11x12")
    for nw in (4, 16):
        sm = synthetic_model(nw, 3, tb_centers=False)
        p = mo.Params(mp_grid=(8, 8, 8), do_ahc=True, do_mep=True, do_kubo=True, do_opt=True, n_hw=5, hw_min=ev(0.05), hw_max=ev(4.0), eta=ev(0.1),
                      valence_bands=nw // 2, T_kelvin=200.0)
        compare(sm, p, f"synthetic nw={nw} 8^3")
    m, p = load_case("femn_d0.900", mp_grid=(64, 64, 64))
    compare(m, p, "femn ahc 64^3 (golden mesh)", threads=os.cpu_count())


if __name__ == "__main__":
    main()
