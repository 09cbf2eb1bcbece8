// mepInterp — C++ host driver of libmepb.so: the reference's response run without Fortran.
//
//   cd <run dir with input.cfg and w90files/<seed>_hr.dat (+ <seed>_r.dat) or <seed>_tb.dat> && mepInterp [--devices 0,1,..] [--check-input]
//
// Mirrors main.f90:24-45 -> init_parameters (src/input_paras.f90:79-293) -> core_worker (src/core.f90:79-298):
// parses ./input.cfg (INI dialect of thirdparty/config_fortran/m_config.f90:226-311), reads the wannier90 _hr.dat as
// read_hr_file does (src/file_io.f90:663-746: degeneracies read and ignored, |x| < 1e-7 eV zeroed, eV -> Hartree), or — like
// read_tb_basis (src/file_io.f90:800-860) — prefers <seed>_tb.dat and otherwise adds <seed>_r.dat when present (the position
// operator switches on the Berry connection / curvature branch),
// runs the whole Monkhorst-Pack mesh through the C ABI (include/mepb.h) and writes out/*.npy + out/occ.dat with the
// names / shapes / Fortran order / units of write_*_tensors (src/file_io.f90:186-398) — NPY v1.0, little endian.
// All physics happens in the CUDA kernels behind mepb_*; there is no CPU compute path in this file.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

#include "../../../include/mepb.h"

namespace {

const double aUtoEv = 27.211385;   // src/constants.f90:18
const double aUtoAngstrm = 0.52917721092;   // src/constants.f90:19

struct Config {
    std::map<std::string, std::vector<std::string>> kv;   // "section%key" -> tokens
    bool has(const std::string &k) const { return kv.count(k) != 0; }
    const std::vector<std::string> &toks(const std::string &k) const {
        auto it = kv.find(k);
        if (it == kv.end()) throw std::runtime_error("input.cfg: missing " + k);
        return it->second;
    }
    double real(const std::string &k, double def) const { return has(k) ? std::stod(toks(k).at(0)) : def; }
    int integer(const std::string &k, int def) const { return has(k) ? std::stoi(toks(k).at(0)) : def; }
    std::string str(const std::string &k, const std::string &def) const { return has(k) ? toks(k).at(0) : def; }
    bool logical(const std::string &k, bool def) const {          // Fortran list-directed logicals: T F .true. .false. True False
        if (!has(k)) return def;
        std::string t = toks(k).at(0);
        size_t i = 0;
        while (i < t.size() && t[i] == '.') ++i;
        if (i >= t.size()) throw std::runtime_error("input.cfg: bad logical for " + k);
        const char ch = (char)std::tolower((unsigned char)t[i]);
        if (ch == 't') return true;
        if (ch == 'f') return false;
        throw std::runtime_error("input.cfg: bad logical for " + k);
    }
    std::vector<double> reals(const std::string &k, size_t n) const {
        const auto &t = toks(k);
        if (t.size() < n) throw std::runtime_error("input.cfg: " + k + " needs " + std::to_string(n) + " values");
        std::vector<double> v(n);
        for (size_t i = 0; i < n; ++i) v[i] = std::stod(t[i]);
        return v;
    }
};

Config parse_cfg(const std::string &path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    Config c;
    std::string line, sec;
    while (std::getline(in, line)) {
        const size_t cpos = line.find_first_of("#;");
        if (cpos != std::string::npos) line.erase(cpos);
        const size_t b = line.find_first_not_of(" \t\r");
        if (b == std::string::npos) continue;
        line = line.substr(b);
        if (line[0] == '[') {
            const size_t e = line.find(']');
            if (e != std::string::npos) {
                sec = line.substr(1, e - 1);
                sec.erase(0, sec.find_first_not_of(" \t")); sec.erase(sec.find_last_not_of(" \t") + 1);
            }
            continue;
        }
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        std::string key = line.substr(0, eq), val = line.substr(eq + 1);
        key.erase(key.find_last_not_of(" \t") + 1);
        for (char &ch : val) if (ch == ',' || ch == '\'' || ch == '"' || ch == '\t' || ch == '\r') ch = ' ';
        std::istringstream ss(val);
        std::vector<std::string> toks;
        for (std::string t; ss >> t;) toks.push_back(t);
        c.kv[sec + "%" + key] = toks;
    }
    return c;
}

struct Inputs {
    mepb_params p{};
    double a_latt[9];                       // Fortran a_latt(3,3): a_latt(i,:) = a_i * a0 -> column-major storage
    std::vector<double> wf_centers;         // (3, N_wf) or empty
    std::string seed;
    bool plot_bands = false, do_write_velo = false;
    int nw = 0, nR = 0;
    std::vector<double> R, H;               // (3,nR), (nw,nw,nR) complex interleaved, Hartree
    std::vector<double> r;                  // (3,nw,nw,nR) complex interleaved, Bohr; empty without a position operator
};

void read_inputs_cfg(const std::string &path, Inputs &I) {
    const Config c = parse_cfg(path);
    const double a0 = c.real("unitCell%a0", 1.0);
    const char *names[3] = {"unitCell%a1", "unitCell%a2", "unitCell%a3"};
    for (int i = 0; i < 3; ++i) {
        const auto v = c.reals(names[i], 3);
        for (int j = 0; j < 3; ++j) I.a_latt[i + 3 * j] = a0 * v[j];          // a_latt(i,j), input_paras.f90:165-168
    }
    const int n_wf = c.integer("wannBase%N_wf", 0);
    if (n_wf > 0) {                                                            // input_paras.f90:127-132
        const auto x = c.reals("wannBase%wf_centers_x", n_wf), y = c.reals("wannBase%wf_centers_y", n_wf),
                   z = c.reals("wannBase%wf_centers_z", n_wf);
        I.wf_centers.resize(3 * (size_t)n_wf);
        for (int n = 0; n < n_wf; ++n) { I.wf_centers[3 * n] = x[n]; I.wf_centers[3 * n + 1] = y[n]; I.wf_centers[3 * n + 2] = z[n]; }
    }
    I.seed = c.str("wannBase%seed_name", "wf1");
    I.plot_bands = c.logical("jobs%plot_bands", false);
    I.do_write_velo = c.logical("jobs%do_write_velo", false);
    mepb_params &p = I.p;
    const auto g = c.reals("wannInterp%mp_grid", 3);
    for (int i = 0; i < 3; ++i) p.mp_grid[i] = (int)g[i];
    p.do_gauge_trafo = c.logical("wannInterp%doGaugeTrafo", true);
    p.use_cart_velo = c.logical("wannInterp%use_cart_velo", false);
    p.do_mep = c.logical("jobs%do_mep", false); p.do_kubo = c.logical("jobs%do_kubo", false);
    p.do_ahc = c.logical("jobs%do_ahc", false); p.do_opt = c.logical("jobs%do_opt", false);
    p.do_write_mep_bands = c.logical("MEP%do_write_mep_bands", false);
    p.valence_bands = c.integer("MEP%valence_bands", 1);
    p.n_hw = std::max(1, c.integer("Laser%N_hw", 1));                          // input_paras.f90:174
    p.n_eF = c.integer("Fermi%N_eF", 1);
    p.hw_min = c.real("Laser%hw_min", 0.0) / aUtoEv; p.hw_max = c.real("Laser%hw_max", 0.0) / aUtoEv;   // :171-172
    p.eF_min = c.real("Fermi%eF_min", 0.0) / aUtoEv; p.eF_max = c.real("Fermi%eF_max", 0.0) / aUtoEv;   // :176-178
    p.T_kelvin = c.real("Fermi%Tkelvin", 0.0);
    p.eta = c.real("Fermi%eta_smearing", 0.0) / aUtoEv;
    p.kubo_tol = c.real("Fermi%kuboTol", 1e-5);
    p.do_gyro = c.logical("jobs%do_gyro", false);
    p.posop_mode = 0;                                                          // PRB 74 195118 form; MEPB_POSOP_MODE=coded selects 1
    if (const char *e = std::getenv("MEPB_POSOP_MODE")) p.posop_mode = (std::string(e) == "coded" || std::string(e) == "1") ? 1 : 0;
}

void read_hr(const std::string &path, Inputs &I) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("[read_hr_file]: cannot open " + path);
    std::string line;
    std::getline(in, line);
    in >> I.nw >> I.nR;
    if (!in || I.nw < 1 || I.nR < 1) throw std::runtime_error("[read_hr_file]: bad header");
    for (int i = 0; i < I.nR; ++i) { int deg; in >> deg; }                      // read and ignored, file_io.f90:685-709
    if (!in) throw std::runtime_error("[read_hr_file]: ERROR issue reading Wigner Seitz degeneracy");
    const size_t nw2 = (size_t)I.nw * I.nw;
    I.R.assign(3 * (size_t)I.nR, 0.0);
    I.H.assign(2 * nw2 * I.nR, 0.0);
    for (int sc = 0; sc < I.nR; ++sc)
        for (size_t wf = 0; wf < nw2; ++wf) {
            double r1, r2, r3, re, im; int m, n;
            in >> r1 >> r2 >> r3 >> m >> n >> re >> im;
            if (!in) throw std::runtime_error("[read_hr_file]: file ends early");
            if (m < 1 || n < 1 || m > I.nw || n > I.nw) throw std::runtime_error("[read_hr_file]: orbital index out of range");
            if (wf == 0) { I.R[3 * sc] = r1; I.R[3 * sc + 1] = r2; I.R[3 * sc + 2] = r3; }   // :723-729
            if (std::fabs(re) < 1e-7) re = 0.0;                                 // :732-733
            if (std::fabs(im) < 1e-7) im = 0.0;
            const size_t idx = 2 * ((size_t)(m - 1) + (size_t)I.nw * (n - 1) + nw2 * sc);
            I.H[idx] = re / aUtoEv; I.H[idx + 1] = im / aUtoEv;
        }
}

// <seed>_r.dat, file_io.f90:750-795: comment; num_wann; nR*nw^2 rows "R1 R2 R3 m n x_re x_im y_re y_im z_re z_im" (Angstrom), R
// blocks in the order of the _hr.dat (checked :772-776).  (The reference mangles the path, :838 vs :760; the format is the one
// scripts/tb_input_writer.py:75-97 writes.)
void read_r(const std::string &path, Inputs &I) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("[read_r_file]: cannot open " + path);
    std::string line;
    std::getline(in, line);
    int f_nw = 0;
    in >> f_nw;
    if (!in || f_nw != I.nw) throw std::runtime_error("[read_tb_basis]: r_mat and H_mat have different basis");
    const size_t nw2 = (size_t)I.nw * I.nw;
    I.r.assign(6 * nw2 * I.nR, 0.0);
    for (int sc = 0; sc < I.nR; ++sc)
        for (size_t it = 0; it < nw2; ++it) {
            double r3[3], v[6]; int m, n;
            in >> r3[0] >> r3[1] >> r3[2] >> m >> n >> v[0] >> v[1] >> v[2] >> v[3] >> v[4] >> v[5];
            if (!in) throw std::runtime_error("[read_r_file]: file ends early");
            for (int i = 0; i < 3; ++i)
                if (std::fabs(r3[i] - I.R[3 * sc + i]) > 1e-8) throw std::runtime_error("[read_r_file]: different R_vect order in _hr.dat and _r.dat file");
            if (m < 1 || n < 1 || m > I.nw || n > I.nw) throw std::runtime_error("[read_r_file]: orbital index out of range");
            for (int a = 0; a < 3; ++a) {
                const size_t idx = 2 * ((size_t)a + 3 * ((size_t)(m - 1) + (size_t)I.nw * (n - 1) + nw2 * sc));
                I.r[idx] = v[2 * a] / aUtoAngstrm; I.r[idx + 1] = v[2 * a + 1] / aUtoAngstrm;
            }
        }
}

// wannier90 <seed>_tb.dat, file_io.f90:555-658: comment; lattice (Angstrom, OVERRIDES a_latt :575-580); num_wann; nrpts; degeneracies
// (skipped); per R: "R1 R2 R3", nw^2 rows "m n Re Im" (eV); then per R: "R1 R2 R3", nw^2 rows "m n x_re x_im y_re y_im z_re z_im"
// (Angstrom).  R kept as the file's integer triples.
void read_tb(const std::string &path, Inputs &I) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("[read_tb_file]: cannot open " + path);
    std::string line;
    std::getline(in, line);
    for (int i = 0; i < 3; ++i) { double v[3]; in >> v[0] >> v[1] >> v[2]; for (int j = 0; j < 3; ++j) I.a_latt[i + 3 * j] = v[j] / aUtoAngstrm; }
    in >> I.nw >> I.nR;
    if (!in || I.nw < 1 || I.nR < 1) throw std::runtime_error("[read_tb_file]: bad header");
    for (int i = 0; i < I.nR; ++i) { int deg; in >> deg; }
    const size_t nw2 = (size_t)I.nw * I.nw;
    I.R.assign(3 * (size_t)I.nR, 0.0); I.H.assign(2 * nw2 * I.nR, 0.0); I.r.assign(6 * nw2 * I.nR, 0.0);
    for (int sc = 0; sc < I.nR; ++sc) {
        in >> I.R[3 * sc] >> I.R[3 * sc + 1] >> I.R[3 * sc + 2];
        for (size_t it = 0; it < nw2; ++it) {
            int m, n; double re, im;
            in >> m >> n >> re >> im;
            if (!in || m < 1 || n < 1 || m > I.nw || n > I.nw) throw std::runtime_error("[read_tb_file]: bad hopping row");
            if (std::fabs(re) < 1e-7) re = 0.0;
            if (std::fabs(im) < 1e-7) im = 0.0;
            const size_t idx = 2 * ((size_t)(m - 1) + (size_t)I.nw * (n - 1) + nw2 * sc);
            I.H[idx] = re / aUtoEv; I.H[idx + 1] = im / aUtoEv;
        }
    }
    for (int sc = 0; sc < I.nR; ++sc) {
        double r3[3];
        in >> r3[0] >> r3[1] >> r3[2];
        for (int i = 0; i < 3; ++i)
            if (std::fabs(r3[i] - I.R[3 * sc + i]) > 1e-8) throw std::runtime_error("[read_tb_basis]: rHopp has different sc order than tHopp");
        for (size_t it = 0; it < nw2; ++it) {
            int m, n; double v[6];
            in >> m >> n >> v[0] >> v[1] >> v[2] >> v[3] >> v[4] >> v[5];
            if (!in || m < 1 || n < 1 || m > I.nw || n > I.nw) throw std::runtime_error("[read_tb_file]: bad position row");
            for (int a = 0; a < 3; ++a) {
                const size_t idx = 2 * ((size_t)a + 3 * ((size_t)(m - 1) + (size_t)I.nw * (n - 1) + nw2 * sc));
                I.r[idx] = v[2 * a] / aUtoAngstrm; I.r[idx + 1] = v[2 * a + 1] / aUtoAngstrm;
            }
        }
    }
}

bool file_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

// NPY v1.0, little endian, fortran_order True (what numpy.load needs to give back the Fortran array of save_npy)
void save_npy(const std::string &path, const double *data, const std::vector<long> &shape, bool cplx) {
    std::string sh = "(";
    size_t count = 1;
    for (size_t i = 0; i < shape.size(); ++i) { sh += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : ""); count *= (size_t)shape[i]; }
    sh += ")";
    std::string hdr = std::string("{'descr': '") + (cplx ? "<c16" : "<f8") + "', 'fortran_order': True, 'shape': " + sh + ", }";
    const size_t pre = 10;
    size_t total = pre + hdr.size() + 1;
    const size_t pad = (64 - total % 64) % 64;
    hdr += std::string(pad, ' ') + "\n";
    std::ofstream out(path, std::ios::binary);
    if (!out) throw std::runtime_error("cannot write " + path);
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    out.write((const char *)magic, 8);
    const unsigned short hl = (unsigned short)hdr.size();
    out.write((const char *)&hl, 2);
    out.write(hdr.data(), (std::streamsize)hdr.size());
    out.write((const char *)data, (std::streamsize)(sizeof(double) * count * (cplx ? 2 : 1)));
}

void check(int rc, const char *what) {
    if (rc != MEPB_OK) throw std::runtime_error(std::string(what) + ": " + mepb_strerror(rc) + " — " + mepb_last_error());
}

}  // namespace

// Band-structure mode (band_calc.f90:30-92): `kpts` (kptsgen.pl format, file_io.f90:451-478: "num_kpts x" then
// "k1 k2 k3 weight" rows, fractional) -> raw/enK.<7-digit ki> (raw stream of nw doubles, file_io.f90:57-72), optionally
// out/velo/velo_Vka.<ki> (file_io.f90:146-185), then out/eBands.dat (file_io.f90:116-143).  One batched mepb_interp call
// replaces the per-k get_wann_interp loop.
int band_mode(const std::string &dir, const Inputs &I, const mepb_model &M, int device) {
    std::ifstream f(dir + "/kpts");
    if (!f) throw std::runtime_error("for bandstructure calculations a kpts file has to be provided");   // band_calc.f90:86
    long long nk = 0; double dummy = 0.0;
    { std::string line; std::getline(f, line); std::stringstream ss(line); if (!(ss >> nk >> dummy) || nk < 1) throw std::runtime_error("kpts: bad header"); }
    std::vector<double> kpt(3 * (size_t)nk);
    for (long long k = 0; k < nk; ++k) {
        std::string line;
        if (!std::getline(f, line)) throw std::runtime_error("kpts: fewer rows than the header announces");
        std::stringstream ss(line);
        if (!(ss >> kpt[3 * k] >> kpt[3 * k + 1] >> kpt[3 * k + 2])) throw std::runtime_error("kpts: bad row " + std::to_string(k + 1));
    }
    mepb_ctx *ctx = nullptr;
    check(mepb_create(&M, &I.p, device, &ctx), "mepb_create");
    const int nw = I.nw;
    std::vector<double> en((size_t)nw * nk), V;
    if (I.do_write_velo) V.resize(2 * 3 * (size_t)nw * nw * nk);
    check(mepb_interp(ctx, nk, kpt.data(), en.data(), I.do_write_velo ? V.data() : nullptr), "mepb_interp");
    mepb_destroy(ctx);
    mkdir((dir + "/out").c_str(), 0755); mkdir((dir + "/raw").c_str(), 0755);
    if (I.do_write_velo) mkdir((dir + "/out/velo").c_str(), 0755);
    char name[64];
    for (long long k = 0; k < nk; ++k) {
        std::snprintf(name, sizeof(name), "/raw/enK.%07lld", k + 1);
        FILE *o = std::fopen((dir + name).c_str(), "wb");
        if (!o) throw std::runtime_error(std::string("cannot write ") + name);
        std::fwrite(&en[(size_t)nw * k], sizeof(double), nw, o);
        std::fclose(o);
        if (I.do_write_velo) {
            std::snprintf(name, sizeof(name), "/out/velo/velo_Vka.%07lld", k + 1);
            FILE *v = std::fopen((dir + name).c_str(), "w");
            if (!v) throw std::runtime_error(std::string("cannot write ") + name);
            std::fprintf(v, " #interpolated velocities\n # x |\t n\t\t|\tm \t|\treal(V^x_nm)  |\t\timag(Vx_nm)\n");
            std::fprintf(v, " #-----------------------------------------------------------------------\n");
            const double *Vk = &V[2 * 3 * (size_t)nw * nw * k];                // V_ka(x,n,m): x + 3 (n + nw m)
            for (int x = 0; x < 3; ++x) {
                for (int m = 0; m < nw; ++m)
                    for (int n = 0; n < nw; ++n) {
                        const double *z = Vk + 2 * (x + 3 * (n + (size_t)nw * m));
                        std::fprintf(v, "\t%1d\t\t%3d\t\t%3d\t\t%16.8f\t%16.8f\n", x + 1, n + 1, m + 1, z[0], z[1]);
                    }
                std::fprintf(v, " #----------------------------------------------------------------------\n");
            }
            std::fclose(v);
        }
    }
    FILE *o = std::fopen((dir + "/out/eBands.dat").c_str(), "w");
    if (!o) throw std::runtime_error("cannot write out/eBands.dat");
    std::fprintf(o, " # energies interpolated by mepinterp-b200 (band mode of MEPinterp)\n");
    std::fprintf(o, " # first 3 columns give the relative k-mesh, 4th column are the enegies\n");
    std::fprintf(o, " # Kpt_idx  K_x (frac)       K_y (frac)        K_z (frac)       Energy (Hartree)  \n");
    for (long long k = 0; k < nk; ++k)
        for (int b = 0; b < nw; ++b)                                           // '(i7,a)' '(200(f16.8))' '(a,f18.8)'
            std::fprintf(o, "%7lld\t%16.8f%16.8f%16.8f\t%18.8f\n", k + 1, kpt[3 * k], kpt[3 * k + 1], kpt[3 * k + 2], en[(size_t)nw * k + b]);
    std::fclose(o);
    std::printf("[mepInterp] band mode: %lld k-points, %d bands; wrote %s/raw/enK.* and %s/out/eBands.dat\n", nk, nw, dir.c_str(), dir.c_str());
    return 0;
}

int main(int argc, char **argv) {
    std::vector<int> devices = {0};
    bool check_only = false;
    std::string dir = ".";
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--check-input") check_only = true;
        else if (a == "--devices" && i + 1 < argc) {
            devices.clear();
            std::stringstream ss(argv[++i]);
            for (std::string t; std::getline(ss, t, ',');) devices.push_back(std::stoi(t));
        } else if (a == "--dir" && i + 1 < argc) dir = argv[++i];
        else { std::fprintf(stderr, "usage: mepInterp [--dir RUN_DIR] [--devices 0,1,..] [--check-input]\n"); return 2; }
    }
    try {
        Inputs I;
        read_inputs_cfg(dir + "/input.cfg", I);
        const std::string seed = dir + "/w90files/" + I.seed;                   // read_tb_basis, file_io.f90:800-860
        if (file_exists(seed + "_tb.dat")) read_tb(seed + "_tb.dat", I);
        else {
            read_hr(seed + "_hr.dat", I);
            if (file_exists(seed + "_r.dat")) read_r(seed + "_r.dat", I);
        }
        if (!I.wf_centers.empty() && (int)(I.wf_centers.size() / 3) != I.nw) throw std::runtime_error("N_wf differs from num_wann");
        const mepb_params &p = I.p;
        if (check_only) {   // machine-readable echo of what was parsed (tests compare it with the oracle's reader)
            std::printf("nw %d nR %d seed %s plot_bands %d\n", I.nw, I.nR, I.seed.c_str(), (int)I.plot_bands);
            {
                double rs = 0.0, ra = 0.0;
                for (size_t i = 0; i < I.r.size(); ++i) { rs += I.r[i] * (double)(i % 11 + 1); ra += std::fabs(I.r[i]); }
                std::printf("posop %d do_gyro %d checksums_r %.17g %.17g\n", (int)!I.r.empty(), p.do_gyro, rs, ra);
            }
            std::printf("mp_grid %d %d %d jobs %d %d %d %d gauge %d cart %d nval %d n_hw %d n_eF %d wmb %d\n", p.mp_grid[0], p.mp_grid[1],
                        p.mp_grid[2], p.do_mep, p.do_kubo, p.do_ahc, p.do_opt, p.do_gauge_trafo, p.use_cart_velo, p.valence_bands, p.n_hw,
                        p.n_eF, p.do_write_mep_bands);
            std::printf("scalars %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", p.hw_min, p.hw_max, p.eF_min, p.eF_max, p.T_kelvin, p.eta, p.kubo_tol);
            std::printf("a_latt"); for (double v : I.a_latt) std::printf(" %.17g", v); std::printf("\n");
            std::printf("wf_centers %zu", I.wf_centers.size()); for (double v : I.wf_centers) std::printf(" %.17g", v); std::printf("\n");
            double sr = 0, sh = 0, sa = 0;
            for (size_t i = 0; i < I.R.size(); ++i) sr += I.R[i] * (double)(i % 7 + 1);
            for (size_t i = 0; i < I.H.size(); ++i) { sh += I.H[i] * (double)(i % 13 + 1); sa += std::fabs(I.H[i]); }
            std::printf("checksums %.17g %.17g %.17g\n", sr, sh, sa);
            return 0;
        }
        mepb_model M{};
        M.nw = I.nw; M.nR = I.nR; M.R = I.R.data(); M.H = I.H.data(); M.r = I.r.empty() ? nullptr : I.r.data();
        M.wf_centers = I.wf_centers.empty() ? nullptr : I.wf_centers.data();
        std::memcpy(M.a_latt, I.a_latt, sizeof(M.a_latt));
        if (I.plot_bands) return band_mode(dir, I, M, devices[0]);            // main.f90:32-36
        const long long nk = (long long)p.mp_grid[0] * p.mp_grid[1] * p.mp_grid[2];
        const int G = (int)devices.size();
        // what `mpirun -np G` is to the reference: a context per device, round-robin blocks of the mesh, ONE NCCL all-reduce of
        // the packed partial sums inside the library (the MPI_REDUCEs of core.f90:467-526), finalize on the first device
        mepb_multi *mm = nullptr;
        check(mepb_create_multi(&M, &p, G, devices.data(), &mm), "mepb_create_multi");
        // outputs
        const int n_hw = p.n_hw, n_eF = p.n_eF, nval = p.valence_bands;
        std::vector<double> Ne(n_eF), mep_ic(9), mep_lc(9), mep_cs(9), mep_bands(9 * (size_t)std::max(nval, 1)), k_ic(9), k_lc(9), k_cs(9), ahc(9),
            velo(18 * (size_t)n_hw), ohc(18 * (size_t)n_hw), opt_s(18 * (size_t)n_hw), opt_a(18 * (size_t)n_hw), photo(54 * (size_t)n_hw * n_eF),
            gyro_c(18 * (size_t)n_eF), gyro_d(18 * (size_t)n_eF), gyro_dw(18 * (size_t)n_hw * n_eF);
        mepb_result r{};
        r.Ne = Ne.data();
        if (p.do_mep) { r.mep_ic = mep_ic.data(); r.mep_lc = mep_lc.data(); r.mep_cs = mep_cs.data(); if (p.do_write_mep_bands) r.mep_bands = mep_bands.data(); }
        if (p.do_kubo) { r.kubo_mep_ic = k_ic.data(); r.kubo_mep_lc = k_lc.data(); r.kubo_mep_cs = k_cs.data(); }
        if (p.do_ahc) { r.ahc = ahc.data(); r.ahc_velo = velo.data(); r.ohc_velo = ohc.data(); }
        if (p.do_opt) { r.opt_s = opt_s.data(); r.opt_a = opt_a.data(); r.photo2 = photo.data(); }
        if (p.do_gyro) { r.gyro_c = gyro_c.data(); r.gyro_d = gyro_d.data(); r.gyro_dw = gyro_dw.data(); }
        check(mepb_multi_run_mesh(mm, &r), "mepb_multi_run_mesh");
        const std::string out = dir + "/out";
        mkdir(out.c_str(), 0755);
        mkdir((dir + "/raw").c_str(), 0755);                                    // input_paras.f90:213-215
        {   // write_hw_list, file_io.f90:186-215: eV when n_hw > 1, raw Hartree otherwise
            std::vector<double> hw(n_hw);
            if (n_hw > 1) for (int i = 0; i < n_hw; ++i) hw[i] = (p.hw_min + i * ((p.hw_max - p.hw_min) / (n_hw - 1))) * aUtoEv;
            else hw[0] = p.hw_min;
            save_npy(out + "/hw_lst.npy", hw.data(), {n_hw}, false);
        }
        if (p.do_mep) {
            std::vector<double> tens(9);
            for (int q = 0; q < 9; ++q) tens[q] = mep_ic[q] + mep_lc[q] + mep_cs[q];
            if (p.do_write_mep_bands) save_npy(out + "/mep_bands.npy", mep_bands.data(), {3, 3, nval}, false);
            save_npy(out + "/mep_ic.npy", mep_ic.data(), {3, 3}, false); save_npy(out + "/mep_lc.npy", mep_lc.data(), {3, 3}, false);
            save_npy(out + "/mep_cs.npy", mep_cs.data(), {3, 3}, false); save_npy(out + "/mep_tens.npy", tens.data(), {3, 3}, false);
        }
        if (p.do_kubo) {
            std::vector<double> tens(9);
            for (int q = 0; q < 9; ++q) tens[q] = k_ic[q] + k_lc[q] + k_cs[q];
            save_npy(out + "/kubo_mep_ic.npy", k_ic.data(), {3, 3}, false); save_npy(out + "/kubo_mep_lc.npy", k_lc.data(), {3, 3}, false);
            save_npy(out + "/kubo_mep_cs.npy", k_cs.data(), {3, 3}, false); save_npy(out + "/kubo_mep_tens.npy", tens.data(), {3, 3}, false);
        }
        if (p.do_ahc) {
            save_npy(out + "/ahc_tens.npy", ahc.data(), {3, 3}, false);
            save_npy(out + "/ahcVELO.npy", velo.data(), {3, 3, n_hw}, true);
            save_npy(out + "/ohcVELO.npy", ohc.data(), {3, 3, n_hw}, true);
        }
        if (p.do_opt) {
            save_npy(out + "/opt_Ssymm.npy", opt_s.data(), {3, 3, n_hw}, true);
            save_npy(out + "/opt_Asymm.npy", opt_a.data(), {3, 3, n_hw}, true);
            save_npy(out + "/photoC_2nd.npy", photo.data(), {3, 3, 3, n_hw, n_eF}, true);
        }
        if (p.do_gyro) {                                                       // write_gyro_tensors, file_io.f90:382-397
            save_npy(out + "/gyro_C.npy", gyro_c.data(), {3, 3, n_eF}, true);
            save_npy(out + "/gyro_D.npy", gyro_d.data(), {3, 3, n_eF}, true);
            save_npy(out + "/gyro_Dw.npy", gyro_dw.data(), {3, 3, n_hw, n_eF}, true);
        }
        {   // out/occ.dat, core.f90:378: '(f16.8,a,f16.8)' with a tab
            FILE *f = std::fopen((out + "/occ.dat").c_str(), "w");
            if (f) {
                for (int i = 0; i < n_eF; ++i) {
                    const double ef = n_eF > 1 ? p.eF_min + i * ((p.eF_max - p.eF_min) / (n_eF - 1)) : p.eF_min;
                    std::fprintf(f, "%16.8f\t%16.8f\n", ef * aUtoEv, Ne[i]);
                }
                std::fclose(f);
            }
        }
        double ms = 0.0, ar_ms = 0.0;
        for (int g = 0; g < G; ++g) {
            double a = 0.0, b = 0.0, f = 0.0;
            mepb_get_timing(mepb_multi_ctx(mm, g), &a, &b, &f);
            ms = std::max(ms, a + b); ar_ms = std::max(ar_ms, b);
        }
        std::printf("[mepInterp] %lld k-points on %d GPU(s): %.3f ms on the device (%.4g k-points/s), eF index %d, degenerate pairs %lld; wrote %s/\n",
                    (long long)r.n_k, G, ms, ms > 0 ? (double)nk / (ms * 1e-3) : 0.0, r.eF_idx + 1, (long long)r.n_degenerate, out.c_str());
        if (G > 1) std::printf("[mepInterp] NCCL all-reduce of the packed partial sums: %.3f ms\n", ar_ms);
        mepb_multi_destroy(mm);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "[mepInterp] ERROR: %s\n", e.what());
        return 1;                                                             // the reference `stop`s
    }
}
