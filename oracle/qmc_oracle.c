/*
 * qmc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the mrbp_qmc VMC/DMC hot path of
 * oarodriguez/PhD-QMCLib (v0.17.0).  It exists so that the CUDA engine in
 * phd_qmclib_b200/ can be checked against something that follows the
 * reference's algorithm line by line and that travels to the GPU box (the
 * reference itself is Python+Numba and only exists in the builder container).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path never does.
 *
 * Parity status: PINNED against the live reference.  The reference's own test
 * suite holds no golden vectors for this path, so the
 * pins are outputs of the reference itself, generated in the builder
 * container by oracle/make_golden.py and committed under tests/golden/.
 *
 * Paths below are relative to /root/reference/src/phd_qmclib/.
 *
 * Documented deviations from the reference (none changes a distribution):
 *   D1  RNG.  The reference draws from Numba's per-thread MT19937
 *       (qmc_base/dmc.py:642,730; qmc_base/jastrow/dmc.py:658-667;
 *       qmc_base/vmc.py:413-415,596,636).  Bitwise replay is impossible
 *       (unseeded worker threads under prange).  The oracle and
 *       the engine share a counter-based Philox4x32-10 convention defined in
 *       this file (section "RNG convention"); every function that consumes
 *       random numbers also has a variant taking them as an explicit array,
 *       which is what is validated against the live reference.
 *   D2  density bin index is clamped to num_bins-1 (the reference writes one
 *       row out of bounds when recast returns exactly L).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QMCO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Parameter block: 25 doubles in the order the reference ships them   */
/* through its own flat arrays (mrbp_qmc/model.py:571-686).            */
/* ------------------------------------------------------------------ */
enum {
    P_V0 = 0, P_R, P_GN, P_NOP, P_L, P_RM, P_VDEF, P_DSEP, P_ZA, P_ZB,
    P_FREE, P_IDEAL,                                   /* model_params */
    O_V0, O_R, O_ZA, O_ZB, O_E0, O_K1, O_KP1,         /* obf_params   */
    T_L, T_RM, T_K2, T_BETA, T_ROFF, T_AM,            /* tbf_params   */
    QMCO_NPARAMS
};

/* Python float floor-mod (numba lowers `%` on floats to this). */
static inline double py_fmod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((r < 0.0) != (b < 0.0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

/* Python float floor-division (numba lowers `//` on floats to this). */
static inline double py_floordiv(double vx, double wx)
{
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0 && ((wx < 0.0) != (mod < 0.0))) div -= 1.0;
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return copysign(0.0, vx / wx);
}

/* qmc_base/utils.py:25-32 */
static inline double sign_(double v) { return copysign(1.0, v); }

/* qmc_base/utils.py:35-51 */
static inline double min_distance(double z_i, double z_j, double sc_size)
{
    double sc_half = 0.5 * sc_size;
    double z_ij = z_i - z_j;
    if (fabs(z_ij) > sc_half)
        return -sc_half + py_fmod(z_ij + sc_half, sc_size);
    return z_ij;
}

/* qmc_base/utils.py:55-66 */
static inline double recast_to_supercell(double z, double z_min, double z_max)
{
    double sc_size = z_max - z_min;
    return z_min + py_fmod(z - z_min, sc_size);
}

/* mrbp_qmc/model.py:404-425 */
static double one_body_func(double z, const double *p)
{
    double v0 = p[O_V0], r = p[O_R], e0 = p[O_E0];
    double k1 = p[O_K1], kp1 = p[O_KP1];
    double z_cell = py_fmod(z, 1.0);
    double z_a = 1 / (1 + r), z_b = r / (1 + r);
    if (z_a < z_cell)
        return cosh(kp1 * (z_cell - 1. + 0.5 * z_b));
    double cf = sqrt(1 + v0 / e0 * pow(sinh(0.5 * sqrt(v0 - e0) * z_b), 2.0));
    return cf * cos(k1 * (z_cell - 0.5 * z_a));
}

/* mrbp_qmc/model.py:429-447 */
static double one_body_func_log_dz(double z, const double *p)
{
    double r = p[O_R], k1 = p[O_K1], kp1 = p[O_KP1];
    double z_cell = py_fmod(z, 1.0);
    double z_a = 1 / (1 + r), z_b = r / (1 + r);
    if (z_a < z_cell)
        return kp1 * tanh(kp1 * (z_cell - 1. + 0.5 * z_b));
    return -k1 * tan(k1 * (z_cell - 0.5 * z_a));
}

/* mrbp_qmc/model.py:451-464 */
static double one_body_func_log_dz2(double z, const double *p)
{
    double v0 = p[O_V0], r = p[O_R], e0 = p[O_E0];
    double z_cell = py_fmod(z, 1.0);
    double z_a = 1 / (1 + r);
    return (z_a < z_cell) ? v0 - e0 : -e0;
}

/* mrbp_qmc/model.py:468-486 */
static double two_body_func(double rz, const double *p)
{
    double sc_size = p[T_L], rm = p[T_RM], k2 = p[T_K2];
    double beta = p[T_BETA], r_off = p[T_ROFF], am = p[T_AM];
    if (rz < fabs(rm))
        return am * cos(k2 * (rz - r_off));
    return pow(sin(M_PI * rz / sc_size), beta);
}

/* mrbp_qmc/model.py:490-507 */
static double two_body_func_log_dz(double rz, const double *p)
{
    double sc_size = p[T_L], rm = p[T_RM], k2 = p[T_K2];
    double beta = p[T_BETA], r_off = p[T_ROFF];
    if (rz < fabs(rm))
        return -k2 * tan(k2 * (rz - r_off));
    return (M_PI / sc_size) * beta / (tan(M_PI * rz / sc_size));
}

/* mrbp_qmc/model.py:511-529 */
static double two_body_func_log_dz2(double rz, const double *p)
{
    double sc_size = p[T_L], rm = p[T_RM], k2 = p[T_K2], beta = p[T_BETA];
    if (rz < fabs(rm))
        return -k2 * k2;
    double t = tan(M_PI * rz / sc_size);
    return pow(M_PI / sc_size, 2.0) * beta * ((beta - 1) / (t * t) - 1);
}

/* mrbp_qmc/model.py:533-551 */
static double potential(double z, const double *p)
{
    double v0 = p[P_V0], z_a = p[P_ZA], v0d = p[P_VDEF];
    double defects_sep = p[P_DSEP];
    double z_cell = py_fmod(z, 1.0);
    double n_cell = floor(z);   /* divmod(z, 1) */
    if (py_fmod(n_cell, defects_sep) == 0.0)
        return (z_a < z_cell) ? v0d : 0.;
    return (z_a < z_cell) ? v0 : 0.;
}

/* qmc_base/jastrow/model.py:287-333 */
static double ith_wf_abs_log(int i, const double *pos, const double *p)
{
    double acc = 0.;
    int nop = (int) p[P_NOP];
    if (!(p[P_FREE] != 0.0)) {
        double obv = one_body_func(pos[i], p);
        acc += log(fabs(obv));
    }
    if (!(p[P_IDEAL] != 0.0)) {
        double z_i = pos[i];
        for (int j = i + 1; j < nop; ++j) {
            double z_ij = min_distance(z_i, pos[j], p[P_L]);
            double tbv = two_body_func(fabs(z_ij), p);
            acc += log(fabs(tbv));
        }
    }
    return acc;
}

/* qmc_base/jastrow/model.py:336-368 */
static double wf_abs_log(const double *pos, const double *p)
{
    double acc = 0.;
    if (p[P_FREE] != 0.0 && p[P_IDEAL] != 0.0) return acc;
    int nop = (int) p[P_NOP];
    for (int i = 0; i < nop; ++i) acc += ith_wf_abs_log(i, pos, p);
    return acc;
}

/* qmc_base/jastrow/model.py:464-525 */
__attribute__((unused)) static double ith_drift(int i, const double *pos, const double *p)
{
    double d = 0.;
    if (p[P_FREE] != 0.0 && p[P_IDEAL] != 0.0) return d;
    double z_i = pos[i];
    if (!(p[P_FREE] != 0.0)) d += one_body_func_log_dz(z_i, p);
    if (!(p[P_IDEAL] != 0.0)) {
        int nop = (int) p[P_NOP];
        for (int j = 0; j < nop; ++j) {
            if (j == i) continue;
            double z_ij = min_distance(z_i, pos[j], p[P_L]);
            double sgn = sign_(z_ij);
            d += two_body_func_log_dz(fabs(z_ij), p) * sgn;
        }
    }
    return d;
}

/* qmc_base/jastrow/model.py:778-856 (and its twin ith_energy, :665-745) */
static void ith_energy_and_drift(int i, const double *pos, const double *p,
                                 double *e_out, double *f_out)
{
    *e_out = 0.; *f_out = 0.;
    if (p[P_FREE] != 0.0 && p[P_IDEAL] != 0.0) return;
    double kin = 0., pot = 0., drift = 0.;
    double z_i = pos[i];
    if (!(p[P_FREE] != 0.0)) {
        double ldz2 = one_body_func_log_dz2(z_i, p);
        double ldz = one_body_func_log_dz(z_i, p);
        kin += (-ldz2 + ldz * ldz);
        pot += potential(z_i, p);
        drift += ldz;
    }
    if (!(p[P_IDEAL] != 0.0)) {
        int nop = (int) p[P_NOP];
        for (int j = 0; j < nop; ++j) {
            if (j == i) continue;
            double z_ij = min_distance(z_i, pos[j], p[P_L]);
            double sgn = sign_(z_ij);
            double ldz2 = two_body_func_log_dz2(fabs(z_ij), p);
            double ldz = two_body_func_log_dz(fabs(z_ij), p) * sgn;
            kin += (-ldz2 + ldz * ldz);
            drift += ldz;
        }
    }
    *e_out = kin - drift * drift + pot;
    *f_out = drift;
}

/* ------------------------------------------------------------------ */
/* Fixed-configuration entry point.                    */
/* confs: [B][2][N] (row 0 positions, row 1 ignored).                  */
/* Any output pointer may be NULL.  drift: [B][N].                     */
/* wf_abs_log -> model.py:336; energy -> :748-775; drift -> :528-566.  */
/* ------------------------------------------------------------------ */
QMCO_API void qmco_model_eval(const double *p, const double *confs,
                              int64_t nconf, double *lnpsi, double *energy,
                              double *drift)
{
    int nop = (int) p[P_NOP];
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nconf; ++b) {
        const double *pos = confs + b * 2 * nop;
        if (lnpsi) lnpsi[b] = wf_abs_log(pos, p);
        if (energy || drift) {
            double e = 0.;
            for (int i = 0; i < nop; ++i) {
                double ei, fi;
                ith_energy_and_drift(i, pos, p, &ei, &fi);
                e += ei;
                /* `ith_drift` (model.py:464-525) sums the same terms in the
                 * same order as ith_energy_and_drift: one pass serves both. */
                if (drift) drift[b * nop + i] = fi;
            }
            if (energy) energy[b] = e;
        }
    }
}

/* qmc_base/jastrow/model.py:968-1004: rho_k = sum cos(k z) + i sum sin(k z) */
static void fourier_density(double kz, const double *pos, int nop,
                            double *re, double *im)
{
    double s_sin = 0., s_cos = 0.;
    for (int i = 0; i < nop; ++i) {
        s_cos += cos(kz * pos[i]);
        s_sin += sin(kz * pos[i]);
    }
    *re = s_cos; *im = s_sin;
}

/* out: [B][M][3] = (|rho_k|^2, Re, Im), k_m = m * 2 pi / L, m = 0..M-1
 * (slot order qmc_base/dmc.py SSFPartSlot; momenta mrbp_qmc/dmc.py:633). */
QMCO_API void qmco_fourier_density(const double *p, const double *confs,
                                   int64_t nconf, int num_modes, double *out)
{
    int nop = (int) p[P_NOP];
    double L = p[P_L];
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nconf; ++b) {
        const double *pos = confs + b * 2 * nop;
        for (int m = 0; m < num_modes; ++m) {
            double kz = m * 2 * M_PI / L;
            double re, im;
            fourier_density(kz, pos, nop, &re, &im);
            double *o = out + (b * num_modes + m) * 3;
            o[0] = re * re + im * im;   /* (a+ib)(a-ib).real */
            o[1] = re;
            o[2] = im;
        }
    }
}

/* The gufunc PhysicalFuncs.fourier_density
 * (qmc_base/jastrow/model.py:1093-1122): arbitrary momenta kz[nk];
 * out: [nconf][nk][2] = (Re, Im). */
QMCO_API void qmco_fourier_density_k(const double *p, const double *confs,
                                     int64_t nconf, const double *kz,
                                     int64_t nk, double *out)
{
    int nop = (int) p[P_NOP];
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nconf; ++b) {
        const double *pos = confs + b * 2 * nop;
        for (int64_t m = 0; m < nk; ++m)
            fourier_density(kz[m], pos, nop, out + (b * nk + m) * 2,
                            out + (b * nk + m) * 2 + 1);
    }
}

/* qmc_base/jastrow/model.py:859-965: local one-body density matrix,
 * <Psi(z_i + sz) / Psi(z_i)> averaged over the particles.
 * out: [nconf][nsz]. */
static double ith_one_body_density(int i, double sz, const double *pos,
                                   const double *p)
{
    int nop = (int) p[P_NOP];
    double acc = 0.;
    if (p[P_FREE] != 0. && p[P_IDEAL] != 0.) return acc;   /* as the reference */
    double z_i = pos[i], z_s = z_i + sz;
    if (p[P_FREE] == 0.)
        acc += log(one_body_func(z_s, p)) - log(one_body_func(z_i, p));
    if (p[P_IDEAL] == 0.) {
        for (int j = 0; j < nop; ++j) {
            if (j == i) continue;
            double d0 = min_distance(z_i, pos[j], p[P_L]);
            double d1 = min_distance(z_s, pos[j], p[P_L]);
            acc += log(two_body_func(fabs(d1), p))
                   - log(two_body_func(fabs(d0), p));
        }
    }
    return exp(acc);
}

QMCO_API void qmco_one_body_density(const double *p, const double *confs,
                                    int64_t nconf, const double *sz,
                                    int64_t nsz, double *out)
{
    int nop = (int) p[P_NOP];
#pragma omp parallel for schedule(static) collapse(2)
    for (int64_t b = 0; b < nconf; ++b)
        for (int64_t s = 0; s < nsz; ++s) {
            const double *pos = confs + b * 2 * nop;
            double obd = 0.;
            for (int i = 0; i < nop; ++i)
                obd += ith_one_body_density(i, sz[s], pos, p);
            out[b * nsz + s] = obd / nop;
        }
}

/* ------------------------------------------------------------------ */
/* RNG convention (deviation D1), shared bit-for-bit with the engine   */
/* (phd_qmclib_b200/csrc/qmcb_rng.cuh).                                */
/*   Philox4x32-10, key = (seed_lo, seed_hi).                          */
/*   counter = (c0, c1, c2, stream):                                   */
/*     stream 0  DMC branching   c0 = global slot, c1 = 0, c2 = step   */
/*     stream 1  DMC diffusion   c0 = global slot of the PARENT,       */
/*               c1 = q | (clone index << 16), c2 = step               */
/*               -> normals for particles 4q .. 4q+3 of the parent's   */
/*               clone-th child (independent of how a multi-GPU run    */
/*               cuts the ensemble into ranks)                         */
/*     stream 2  VMC proposal    c0 = chain, c1 = q, c2 = step         */
/*               -> uniforms / normals for particles 4q .. 4q+3        */
/*     stream 3  VMC acceptance  c0 = chain, c1 = 0, c2 = step         */
/*   streams 0, 3 (one draw per call, 53 bits):                        */
/*     uniform  u = ((x0 << 21) ^ (x1 >> 11)) * 2^-53      in [0, 1)   */
/*   streams 1, 2 (four draws per call, 32 bits each):                 */
/*     uniform  u_i = (x_i + 1/2) 2^-32                    in (0, 1)   */
/*     normals  (Box-Muller) r_a = sqrt(-2 ln((x0 + 1/2) 2^-32)),      */
/*              (n0, n1) = r_a (cos, sin)(2 pi x1 2^-32),              */
/*              (n2, n3) likewise from x2, x3                          */
/* ------------------------------------------------------------------ */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t) 0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t) p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

static inline double u53(uint32_t a, uint32_t b)
{
    uint64_t m = (((uint64_t) a) << 21) ^ (((uint64_t) b) >> 11);
    return (double) m * (1.0 / 9007199254740992.0);
}

static inline void rng_uniform2(uint64_t seed, uint32_t c0, uint32_t c1,
                                uint32_t c2, uint32_t stream, double *u0,
                                double *u1)
{
    uint32_t c[4] = {c0, c1, c2, stream};
    philox4x32_10(c, (uint32_t) seed, (uint32_t) (seed >> 32));
    *u0 = u53(c[0], c[1]);
    *u1 = u53(c[2], c[3]);
}

static inline double u32_open(uint32_t x)
{
    return ((double) x + 0.5) * (1.0 / 4294967296.0);
}

static inline void rng_uniform4(uint64_t seed, uint32_t c0, uint32_t c1,
                                uint32_t c2, uint32_t stream, double u[4])
{
    uint32_t c[4] = {c0, c1, c2, stream};
    philox4x32_10(c, (uint32_t) seed, (uint32_t) (seed >> 32));
    for (int i = 0; i < 4; ++i) u[i] = u32_open(c[i]);
}

static inline void rng_normal4(uint64_t seed, uint32_t c0, uint32_t c1,
                               uint32_t c2, uint32_t stream, double n[4])
{
    uint32_t c[4] = {c0, c1, c2, stream};
    philox4x32_10(c, (uint32_t) seed, (uint32_t) (seed >> 32));
    for (int h = 0; h < 2; ++h) {
        double r = sqrt(-2.0 * log(u32_open(c[2 * h])));
        double th = 2.0 * M_PI * ((double) c[2 * h + 1] * (1.0 / 4294967296.0));
        n[2 * h] = r * cos(th);
        n[2 * h + 1] = r * sin(th);
    }
}

/* Expose the raw streams so tests can feed identical numbers elsewhere. */
QMCO_API void qmco_rng_uniform2(uint64_t seed, uint32_t c0, uint32_t c1,
                                uint32_t c2, uint32_t stream, double *out)
{
    rng_uniform2(seed, c0, c1, c2, stream, out, out + 1);
}

QMCO_API void qmco_rng_normal4(uint64_t seed, uint32_t c0, uint32_t c1,
                               uint32_t c2, uint32_t stream, double *out)
{
    rng_normal4(seed, c0, c1, c2, stream, out);
}

QMCO_API void qmco_rng_uniform4(uint64_t seed, uint32_t c0, uint32_t c1,
                                uint32_t c2, uint32_t stream, double *out)
{
    rng_uniform4(seed, c0, c1, c2, stream, out);
}

/* ------------------------------------------------------------------ */
/* DMC                                                                 */
/* ------------------------------------------------------------------ */

/* qmc_base/dmc.py:614-655.  uniforms[s] replaces random.rand() for parent s.
 * Returns the new number of walkers; ref[0..W) = parent slot. */
QMCO_API int64_t qmco_branch(const double *weights, int64_t prev_num_walkers,
                             int64_t max_num_walkers, const double *uniforms,
                             int64_t *cloning_ref)
{
    int64_t final_num = 0;
    for (int64_t s = 0; s < prev_num_walkers; ++s) {
        if (final_num >= max_num_walkers) break;
        int64_t clone_factor = (int64_t) (weights[s] + uniforms[s]);
        if (!clone_factor) continue;
        int64_t start = final_num;
        final_num = final_num + clone_factor;
        if (final_num > max_num_walkers) final_num = max_num_walkers;
        for (int64_t c = start; c < final_num; ++c) cloning_ref[c] = s;
    }
    return final_num;
}

/* One call of evolve_state (qmc_base/jastrow/dmc.py:830-951 with
 * evolve_system :743-827 and ith_diffusion :634-673, recast
 * mrbp_qmc/dmc.py:453-469).
 *
 * State buffers (prev / act / next), each: confs [Wmax][2][N], energy[Wmax],
 * weight[Wmax], mask[Wmax] (uint8).  normals: [Wmax][N], the N(0, sigma)
 * draws for slot s (already scaled by sigma = sqrt(2 dt)).
 * energy_mode 0 = reference semantics (stale slot energy, quirk Q1);
 *             1 = textbook (parent's energy). */
QMCO_API void qmco_evolve_state(const double *p,
                                const double *prev_confs,
                                const double *prev_energy,
                                double *act_confs, double *act_energy,
                                double *act_weight, uint8_t *act_mask,
                                double *next_confs, double *next_energy,
                                double *next_weight,
                                int64_t num_walkers, int64_t max_num_walkers,
                                double time_step, double ref_energy,
                                const int64_t *cloning_ref,
                                const double *normals,
                                double z_min, double z_max, int energy_mode)
{
    int nop = (int) p[P_NOP];
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t s = 0; s < max_num_walkers; ++s) {
        if (s >= num_walkers) {
            act_mask[s] = 1;
            continue;
        }
        int64_t r = cloning_ref[s];
        const double *pc = prev_confs + r * 2 * nop;
        double *sc = act_confs + s * 2 * nop;
        double *nc = next_confs + s * 2 * nop;
        double sys_energy = prev_energy[r];

        for (int i = 0; i < nop; ++i) {
            double z_i = pc[i], drift_i = pc[nop + i];
            double z_next = z_i + 2 * drift_i * time_step
                            + normals[s * nop + i];
            double z_rc = recast_to_supercell(z_next, z_min, z_max);
            sc[i] = z_rc;
            nc[i] = z_rc;
        }
        double energy = (energy_mode == 0) ? act_energy[s] : sys_energy;
        double energy_next = 0.;
        for (int i = 0; i < nop; ++i) {
            double ei, fi;
            ith_energy_and_drift(i, sc, p, &ei, &fi);
            nc[nop + i] = fi;
            energy_next += ei;
        }
        double mean_energy = (energy_next + energy) / 2;
        double weight_next = exp(-time_step * (mean_energy - ref_energy));
        next_energy[s] = energy_next;
        next_weight[s] = weight_next;

        /* cloning: the yielded ("actual") state is the parent's data */
        memcpy(sc, pc, sizeof(double) * 2 * nop);
        act_energy[s] = sys_energy;
        act_weight[s] = 1.;
        act_mask[s] = 0;
    }
}

/* prepare_state_data (qmc_base/jastrow/dmc.py:1030-1174) for n configs. */
QMCO_API void qmco_prepare_state(const double *p, const double *ini_confs,
                                 int64_t n, int64_t max_num_walkers,
                                 double *confs, double *energy, double *weight,
                                 uint8_t *mask)
{
    int nop = (int) p[P_NOP];
    for (int64_t s = 0; s < max_num_walkers; ++s) mask[s] = 1;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n; ++s) {
        const double *ic = ini_confs + s * 2 * nop;
        double *sc = confs + s * 2 * nop;
        double e = 0.;
        for (int i = 0; i < nop; ++i) {
            double ei, fi;
            ith_energy_and_drift(i, ic, p, &ei, &fi);
            sc[i] = ic[i];
            sc[nop + i] = fi;
            e += ei;
        }
        energy[s] = e;
        weight[s] = 1.;
        mask[s] = 0;
    }
}

/* S(k) estimator step: fourier_density_inner + _core
 * (qmc_base/jastrow/dmc.py:363-573).  aux: [2][Wmax][M][3]; iter: [nts][M][3].
 */
QMCO_API void qmco_ssf_step(const double *p, int64_t step_idx,
                            const double *confs, int64_t num_walkers,
                            int64_t max_num_walkers,
                            const int64_t *cloning_ref, int num_modes,
                            int as_pure_est, int64_t pfw_nts,
                            double *iter_ssf, double *aux)
{
    int nop = (int) p[P_NOP];
    double L = p[P_L];
    int64_t cur = step_idx % 2, prv = 1 - cur;
    double *a_cur = aux + cur * max_num_walkers * num_modes * 3;
    const double *a_prv = aux + prv * max_num_walkers * num_modes * 3;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < num_walkers; ++s) {
        const double *pos = confs + s * 2 * nop;
        double *mine = a_cur + s * num_modes * 3;
        const double *par = a_prv + cloning_ref[s] * num_modes * 3;
        if (as_pure_est && step_idx >= pfw_nts) {
            memcpy(mine, par, sizeof(double) * num_modes * 3);
            continue;
        }
        for (int m = 0; m < num_modes; ++m) {
            double kz = m * 2 * M_PI / L, re, im;
            fourier_density(kz, pos, nop, &re, &im);
            double sq = re * re + im * im;
            if (!as_pure_est) {
                mine[m * 3 + 0] = sq;
                mine[m * 3 + 1] = re;
                mine[m * 3 + 2] = im;
            } else {
                mine[m * 3 + 0] = sq + par[m * 3 + 0];
                mine[m * 3 + 1] = re + par[m * 3 + 1];
                mine[m * 3 + 2] = im + par[m * 3 + 2];
            }
        }
    }
    double div = 1.;
    if (as_pure_est) div = (step_idx < pfw_nts) ? (double) (step_idx + 1)
                                                : (double) pfw_nts;
    double *it = iter_ssf + step_idx * num_modes * 3;
    for (int m = 0; m < num_modes; ++m)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.;
            for (int64_t s = 0; s < num_walkers; ++s)
                acc += a_cur[(s * num_modes + m) * 3 + c];
            if (as_pure_est) acc /= div;
            it[m * 3 + c] = acc;
        }
}

/* Density estimator step: density_inner (qmc_base/jastrow/dmc.py:195-302) +
 * density_core (mrbp_qmc/dmc.py:472-547).  aux: [2][Wmax][B]; iter: [nts][B].
 * Pure mode copies ALL slots from the other ping-pong buffer and ignores the
 * cloning table (quirk Q2); mixed mode never resets (quirk Q3). */
QMCO_API void qmco_density_step(const double *p, int64_t step_idx,
                                const double *confs, int64_t num_walkers,
                                int64_t max_num_walkers, int num_bins,
                                int as_pure_est, int64_t pfw_nts,
                                double *iter_density, double *aux)
{
    int nop = (int) p[P_NOP];
    double bin_size = p[P_L] / num_bins;
    int64_t cur = step_idx % 2, prv = 1 - cur;
    double *a_cur = aux + cur * max_num_walkers * num_bins;
    const double *a_prv = aux + prv * max_num_walkers * num_bins;
    if (as_pure_est)
        memcpy(a_cur, a_prv, sizeof(double) * max_num_walkers * num_bins);
    if (!as_pure_est || step_idx < pfw_nts) {
#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < num_walkers; ++s) {
            const double *pos = confs + s * 2 * nop;
            double *mine = a_cur + s * num_bins;
            for (int i = 0; i < nop; ++i) {
                int64_t b = (int64_t) py_floordiv(pos[i], bin_size);
                if (b >= num_bins) b = num_bins - 1;      /* D2 */
                if (b < 0) b = 0;
                mine[b] += 1;
            }
        }
    }
    double div = 1.;
    if (as_pure_est) div = (step_idx < pfw_nts) ? (double) (step_idx + 1)
                                                : (double) pfw_nts;
    double *it = iter_density + step_idx * num_bins;
    for (int b = 0; b < num_bins; ++b) {
        double acc = 0.;
        for (int64_t s = 0; s < num_walkers; ++s) acc += a_cur[s * num_bins + b];
        if (as_pure_est) acc /= div;
        it[b] = acc;
    }
}

/* Whole-run driver: states_generator (qmc_base/dmc.py:664-787) wrapped by
 * blocks (:815-971) for ONE block of nts steps, RNG per convention D1.
 * The three state buffers are owned by the caller and persist between calls
 * (prev / act / next as in the reference; this function swaps prev and next
 * internally and reports which is which through *parity).
 *
 * scal: [0]=ref_energy [1]=total_energy [2]=total_weight (in/out)
 * cnt : [0]=prev_num_walkers [1]=global step counter (in/out)
 */
typedef struct {
    double *confs, *energy, *weight;
    uint8_t *mask;
} qmco_state_buf;

QMCO_API void qmco_dmc_block(const double *p, uint64_t seed,
                             double time_step, int64_t target_num_walkers,
                             double nwc_factor, int64_t max_num_walkers,
                             double z_min, double z_max, int energy_mode,
                             qmco_state_buf *buf_a /* prev */,
                             qmco_state_buf *buf_act,
                             qmco_state_buf *buf_b /* next */,
                             int64_t *cloning_ref, double *scal, int64_t *cnt,
                             int64_t nts,
                             double *it_energy, double *it_weight,
                             uint64_t *it_num_walkers, double *it_ref_energy,
                             double *it_accum_energy,
                             /* estimators, evaluated iff eval_est != 0 */
                             int eval_est,
                             int ssf_modes, int ssf_pure, int64_t ssf_pfw,
                             double *iter_ssf, double *aux_ssf,
                             int dens_bins, int dens_pure, int64_t dens_pfw,
                             double *iter_density, double *aux_density,
                             /* optional explicit draws replacing Philox:
                              * uniforms_ext [nts][Wmax]    U[0,1) per parent
                              * normals_ext  [nts][Wmax][N] N(0,1) per slot */
                             const double *uniforms_ext,
                             const double *normals_ext)
{
    int nop = (int) p[P_NOP];
    double sigma = sqrt(2 * time_step);
    double ref_energy = scal[0], total_energy = scal[1], total_weight = scal[2];
    int64_t prev_num = cnt[0], gstep = cnt[1];
    qmco_state_buf *prev = buf_a, *next = buf_b;
    double *normals = (double *) malloc(sizeof(double) * max_num_walkers
                                        * (nop + (nop & 1)));
    double *unif = (double *) malloc(sizeof(double) * max_num_walkers);

    for (int64_t step = 0; step < nts; ++step, ++gstep) {
        for (int64_t s = 0; s < prev_num; ++s) {
            double u0, u1;
            if (uniforms_ext) {
                u0 = uniforms_ext[step * max_num_walkers + s];
            } else {
                rng_uniform2(seed, (uint32_t) s, 0u, (uint32_t) gstep, 0u,
                             &u0, &u1);
            }
            unif[s] = u0;
        }
        int64_t nw = qmco_branch(prev->weight, prev_num, max_num_walkers,
                                 unif, cloning_ref);
#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < nw; ++s) {
            int64_t clone = 0;
            while (clone < 0xffff && s - clone - 1 >= 0
                   && cloning_ref[s - clone - 1] == cloning_ref[s])
                ++clone;
            for (int q = 0; 4 * q < nop; ++q) {
                double n4[4] = {0., 0., 0., 0.};
                if (normals_ext) {
                    const double *ne = normals_ext
                        + (step * max_num_walkers + s) * nop;
                    for (int i = 0; i < 4 && 4 * q + i < nop; ++i)
                        n4[i] = ne[4 * q + i];
                } else {
                    rng_normal4(seed, (uint32_t) cloning_ref[s],
                                (uint32_t) q | ((uint32_t) clone << 16),
                                (uint32_t) gstep, 1u, n4);
                }
                for (int i = 0; i < 4 && 4 * q + i < nop; ++i)
                    normals[s * nop + 4 * q + i] = sigma * n4[i];
            }
        }
        qmco_evolve_state(p, prev->confs, prev->energy,
                          buf_act->confs, buf_act->energy, buf_act->weight,
                          buf_act->mask, next->confs, next->energy,
                          next->weight, nw, max_num_walkers, time_step,
                          ref_energy, cloning_ref, normals, z_min, z_max,
                          energy_mode);
        /* qmc_base/dmc.py:758-771 */
        double state_energy = 0., state_weight = 0.;
        for (int64_t s = 0; s < nw; ++s) state_energy += buf_act->energy[s];
        for (int64_t s = 0; s < nw; ++s) state_weight += buf_act->weight[s];
        total_energy += state_energy;
        total_weight += state_weight;
        double accum = total_energy / total_weight;
        ref_energy = accum - nwc_factor
                     * log(state_weight / target_num_walkers) / time_step;

        it_energy[step] = state_energy;
        it_weight[step] = state_weight;
        it_num_walkers[step] = (uint64_t) nw;
        it_ref_energy[step] = ref_energy;
        it_accum_energy[step] = accum;

        if (eval_est) {
            if (dens_bins > 0)
                qmco_density_step(p, step, buf_act->confs, nw,
                                  max_num_walkers, dens_bins, dens_pure,
                                  dens_pfw, iter_density, aux_density);
            if (ssf_modes > 0)
                qmco_ssf_step(p, step, buf_act->confs, nw, max_num_walkers,
                              cloning_ref, ssf_modes, ssf_pure, ssf_pfw,
                              iter_ssf, aux_ssf);
        }
        qmco_state_buf *t = prev; prev = next; next = t;
        prev_num = nw;
    }
    /* hand the (possibly swapped) roles back */
    if (prev != buf_a) {
        qmco_state_buf t = *buf_a; *buf_a = *buf_b; *buf_b = t;
    }
    scal[0] = ref_energy; scal[1] = total_energy; scal[2] = total_weight;
    cnt[0] = prev_num; cnt[1] = gstep;
    free(normals); free(unif);
}

/* ------------------------------------------------------------------ */
/* VMC: states_generator (qmc_base/vmc.py:557-648) + blocks (:670-770) */
/* with jastrow/vmc.py:201-351 and mrbp_qmc/vmc.py:206-271.            */
/* Batched over independent chains (each chain is the reference's      */
/* single-chain algorithm; chain c uses RNG counter c0 = c).           */
/*                                                                     */
/* cur: [C][2][N] current configurations (in/out)                      */
/* lnpsi_cur[C], energy_prev[C], ssf_prev[C][M][3] carry the "previous */
/* entry" the reference copies on rejection (jastrow/vmc.py:252-255).  */
/* first != 0: the first yielded state of the chain is the initial one */
/* flagged ACCEPTED (qmc_base/vmc.py:616-618).                         */
/* outputs per chain: [C][ns] (+[M][3] for ssf); ssf may be NULL.      */
/* uniforms_ext, if not NULL: [ns][C][N+1] explicit draws (N proposal  */
/* draws then the U[0,1) acceptance one) used instead of Philox.       */
/* proposal: 0 = uniform, z + (u - 1/2) move_spread (qmc_base/vmc.py:  */
/* 401-415); 1 = gaussian, z + N(0, sigma = move_spread)               */
/* (qmc_base/vmc_ndf.py:44-62; explicit draws are then N(0,1)).        */
/* ------------------------------------------------------------------ */
QMCO_API void qmco_vmc_block(const double *p, uint64_t seed,
                             double move_spread, double z_min, double z_max,
                             int64_t num_chains, int64_t chain_offset,
                             int64_t ns, int64_t step0, int first,
                             double *cur, double *lnpsi_cur,
                             double *energy_prev, double *ssf_prev,
                             int num_modes,
                             double *out_lnpsi, double *out_energy,
                             uint8_t *out_stat, double *out_ssf,
                             double *accept_rate,
                             const double *uniforms_ext, int proposal)
{
    int nop = (int) p[P_NOP];
    double L = p[P_L];
#pragma omp parallel
    {
        double *prop = (double *) malloc(sizeof(double) * 2 * nop);
#pragma omp for schedule(static)
        for (int64_t c = 0; c < num_chains; ++c) {
            double *cc = cur + c * 2 * nop;
            double ln_cur = lnpsi_cur[c];
            double accepted = 0.;
            for (int64_t st = 0; st < ns; ++st) {
                int stat;
                if (first && st == 0) {
                    stat = 1;
                } else {
                    /* generator step index: the first yield consumes no RNG */
                    int64_t g = step0 + st - (first ? 1 : 0);
                    const double *ue = uniforms_ext
                        ? uniforms_ext + ((st - (first ? 1 : 0)) * num_chains
                                          + c) * (nop + 1)
                        : NULL;
                    for (int q = 0; 4 * q < nop; ++q) {
                        double u4[4] = {0., 0., 0., 0.};
                        if (ue) {
                            for (int i = 0; i < 4 && 4 * q + i < nop; ++i)
                                u4[i] = ue[4 * q + i];
                        } else if (proposal == 1) {
                            rng_normal4(seed, (uint32_t) (chain_offset + c),
                                        (uint32_t) q, (uint32_t) g, 2u, u4);
                        } else {
                            rng_uniform4(seed, (uint32_t) (chain_offset + c),
                                         (uint32_t) q, (uint32_t) g, 2u, u4);
                        }
                        for (int i = 0; i < 4 && 4 * q + i < nop; ++i) {
                            double d = proposal == 1
                                ? move_spread * u4[i]
                                : (u4[i] - 0.5) * move_spread;
                            prop[4 * q + i] = recast_to_supercell(
                                cc[4 * q + i] + d, z_min, z_max);
                        }
                    }
                    double ln_next = wf_abs_log(prop, p);
                    double ua, ub;
                    if (ue) ua = ue[nop];
                    else rng_uniform2(seed, (uint32_t) (chain_offset + c), 0u,
                                      (uint32_t) g, 3u, &ua, &ub);
                    stat = 0;
                    if (ln_next > 0.5 * log(ua) + ln_cur) {
                        memcpy(cc, prop, sizeof(double) * nop);
                        ln_cur = ln_next;
                        stat = 1;
                    }
                }
                accepted += stat;
                out_lnpsi[c * ns + st] = ln_cur;
                out_stat[c * ns + st] = (uint8_t) stat;
                if (stat) {
                    double e = 0.;
                    for (int i = 0; i < nop; ++i) {
                        double ei, fi;
                        ith_energy_and_drift(i, cc, p, &ei, &fi);
                        e += ei;
                    }
                    energy_prev[c] = e;
                }
                out_energy[c * ns + st] = energy_prev[c];
                if (out_ssf) {
                    double *sp = ssf_prev + c * num_modes * 3;
                    if (stat) {
                        for (int m = 0; m < num_modes; ++m) {
                            double re, im;
                            fourier_density(m * 2 * M_PI / L, cc, nop,
                                            &re, &im);
                            sp[m * 3 + 0] = re * re + im * im;
                            sp[m * 3 + 1] = re;
                            sp[m * 3 + 2] = im;
                        }
                    }
                    memcpy(out_ssf + (c * ns + st) * num_modes * 3, sp,
                           sizeof(double) * num_modes * 3);
                }
            }
            lnpsi_cur[c] = ln_cur;
            accept_rate[c] = accepted / (double) ns;
        }
        free(prop);
    }
}

QMCO_API int qmco_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

QMCO_API void qmco_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}
