// Device-side building blocks of the B200 walker-ensemble engine.
//
// Reference formulas (paths relative to src/phd_qmclib/ of PhD-QMCLib):
//   one-body f1 and log-derivatives     mrbp_qmc/model.py:404-464
//   two-body f2 and log-derivatives     mrbp_qmc/model.py:468-529
//   KP potential with defects           mrbp_qmc/model.py:533-551
//   minimum image                       qmc_base/utils.py:35-51
//   lnPsi / drift / local energy        qmc_base/jastrow/model.py:287-368,
//                                       464-566, 665-856
//
// The arithmetic is NOT a transliteration.  With a_i = pi z_i / L and
// u_i = k2 z_i tabulated once per particle (sin, cos), every pair term follows
// from angle-addition identities, one reciprocal and no transcendental:
//   far  (r >= r_m):  f2'/f2 * sgn = (pi/L) beta  cos(a_i-a_j)/sin(a_i-a_j)
//                     -f2''/f2 + (f2'/f2)^2 = (pi/L)^2 beta / sin^2(a_i-a_j)
//   near (r <  r_m):  theta = k2 r - k2 r_off,
//                     f2'/f2 = -k2 tan(theta),
//                     -f2''/f2 + (f2'/f2)^2 = k2^2 / cos^2(theta)
// cot and 1/sin^2 have period L in z_i - z_j, so the far branch needs no
// minimum-image step; the near branch folds the wrap and sgn(d) into a
// rotation by a constant angle.  r < r_m <=> |sin(a_i - a_j)| < sin(pi r_m/L)
// because r lies in [0, L/2].  Results agree with the reference to ~1e-15
// relative (tests/), far inside the 1e-12 budget of the north star.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmcb {

constexpr int TB = 4;                 // particles owned by one thread
constexpr double LN2 = 0.693147180559945309417232121458;

struct DevModel {
    int nop;            // N
    int nb;             // ceil(N / 4): particle blocks == threads per walker
    int kmax;           // circulant half-width: nb / 2
    int is_free, is_ideal;
    int defects_sep;
    double L, inv_L;
    // one-body (Kronig-Penney cell: well [0, za], barrier (za, 1))
    double za, zb, k1, kp1, e0, v0, vdef, ln_cf;
    // two-body
    double s_m;                       // sin(pi r_m / L)
    double beta, ln_am, k2;
    double A_far, B_far, A_near, B_near;
    double cps0, sps0;                // cos/sin(k2 r_off)
    double cps1, sps1;                // cos/sin(k2 r_off - k2 L)
};

// Launch geometry shared by every walker-group kernel.
struct GroupGeom {
    int nthreads;       // CTA size (multiple of 32)
    int G;              // walkers per CTA
    int nbp;            // padded row length of the shared tables (>= nb)
    int smem_bytes;
};

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double flip_sign(double x, int signbit)
{
    return __hiloint2double(__double2hiint(x) ^ signbit, __double2loint(x));
}

// 1/x to <= ~1 ulp for normal x: MUFU.RCP64H seed (rel. err <= 2^-20 or
// better) + one cubically convergent step, 3 DFMA.
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    double p = fma(e, e, e);
    return fma(y, p, y);
}

// Pull the binary exponent of p (> 0) into e, leaving p in [1, 2).
__device__ __forceinline__ void renorm(double &p, int &e)
{
    int hi = __double2hiint(p);
    int ex = (hi >> 20) & 0x7ff;
    if (ex != 0) {
        e += ex - 1023;
        p = __hiloint2double((hi & 0x800fffff) | (1023 << 20),
                             __double2loint(p));
    }
}

// z_min + ((z - z_min) floor-mod (z_max - z_min)); qmc_base/utils.py:55-66.
__device__ __forceinline__ double recast(double z, double z_min, double size)
{
    double x = z - z_min;
    if (x >= 0.0 && x < size) return z_min + x;
    double r;
    if (x < 0.0 && x >= -size) {
        r = x + size;                       // Python: fmod(x) = x, then += b
    } else if (x >= size && x < 2.0 * size) {
        r = x - size;                       // exact
    } else {
        r = fmod(x, size);
        if (r < 0.0) r += size;
    }
    return z_min + r;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 and the RNG convention shared with oracle/qmc_oracle.c
// ---------------------------------------------------------------------------
enum : uint32_t { STREAM_BRANCH = 0, STREAM_DIFFUSE = 1, STREAM_VMC_MOVE = 2,
                  STREAM_VMC_ACCEPT = 3 };

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0,
                                              uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]);
        uint32_t lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]);
        uint32_t lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0;
        uint32_t n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    unsigned long long m = (((unsigned long long) a) << 21)
                           ^ (((unsigned long long) b) >> 11);
    return (double) m * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void rng_uniform2(uint64_t seed, uint32_t c0,
                                             uint32_t c1, uint32_t c2,
                                             uint32_t stream, double &u0,
                                             double &u1)
{
    uint32_t c[4] = {c0, c1, c2, stream};
    philox4x32_10(c, (uint32_t) seed, (uint32_t) (seed >> 32));
    u0 = u53(c[0], c[1]);
    u1 = u53(c[2], c[3]);
}

__device__ __forceinline__ void rng_normal2(uint64_t seed, uint32_t c0,
                                            uint32_t c1, uint32_t c2,
                                            uint32_t stream, double &n0,
                                            double &n1)
{
    double u0, u1;
    rng_uniform2(seed, c0, c1, c2, stream, u0, u1);
    double r = sqrt(-2.0 * log(u0 + 1.0 / 9007199254740992.0));
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    n0 = r * c;
    n1 = r * s;
}

// ---------------------------------------------------------------------------
// one-body terms of one particle
// ---------------------------------------------------------------------------
struct OneBody {
    double ldz;     // f1'/f1
    double kin;     // -f1''/f1 + (f1'/f1)^2
    double pot;     // V(z)
    double lnf;     // ln |f1|
};

template <bool LN>
__device__ __forceinline__ OneBody one_body(const DevModel &M, double z)
{
    OneBody o;
    double n_cell = floor(z);
    double zc = z - n_cell;                         // z mod 1, exact
    if (M.za < zc) {                                // barrier
        double arg = M.kp1 * (zc - 1.0 + 0.5 * M.zb);
        double th = tanh(arg);
        o.ldz = M.kp1 * th;
        o.kin = -(M.v0 - M.e0) + o.ldz * o.ldz;
        bool defect = (M.defects_sep == 1)
                      || (fmod(n_cell, (double) M.defects_sep) == 0.0);
        o.pot = defect ? M.vdef : M.v0;
        if (LN) o.lnf = log(cosh(arg));
    } else {                                        // well
        double arg = M.k1 * (zc - 0.5 * M.za);
        double s, c;
        sincos(arg, &s, &c);
        o.ldz = -M.k1 * (s / c);
        o.kin = M.e0 + o.ldz * o.ldz;
        o.pot = 0.0;
        if (LN) o.lnf = M.ln_cf + log(fabs(c));
    }
    return o;
}

// sin/cos tables of one particle
__device__ __forceinline__ void particle_tables(const DevModel &M, double z,
                                                double &sa, double &ca,
                                                double &su, double &cu)
{
    sincospi(z / M.L, &sa, &ca);
    sincos(M.k2 * z, &su, &cu);
}

// ---------------------------------------------------------------------------
// The pair term.  (i) is the row particle (this thread), (j) the column one.
//   v  : contribution to F_i (and -v to F_j)
//   kk : -f2''/f2 + (f2'/f2)^2 of the unordered pair
//   far/near factor for ln|f2|: |sin(a_i-a_j)| (to the power beta) or
//   |cos(theta)| (times a_m).
// ---------------------------------------------------------------------------
template <bool LN>
__device__ __forceinline__ void pair_term(const DevModel &M,
                                          double sa_i, double ca_i,
                                          double su_i, double cu_i,
                                          double sa_j, double ca_j,
                                          double su_j, double cu_j,
                                          double &v, double &kk,
                                          double &ffar, double &fnear,
                                          int &is_near)
{
    double Sa = fma(sa_i, ca_j, -(ca_i * sa_j));
    double Ca = fma(ca_i, ca_j, sa_i * sa_j);
    double Su = fma(su_i, cu_j, -(cu_i * su_j));
    double Cu = fma(cu_i, cu_j, su_i * su_j);
    bool near = fabs(Sa) < M.s_m;
    int hiC = __double2hiint(Ca);
    bool wrapped = hiC < 0;
    int sgn = (__double2hiint(Sa) ^ hiC) & 0x80000000;       // sign(d_min)
    double sSu = flip_sign(Su, sgn);
    double cps = wrapped ? M.cps1 : M.cps0;
    double sps = wrapped ? M.sps1 : M.sps0;
    double S2 = fma(sSu, cps, -(Cu * sps));                  // sin(theta)
    double C2 = fma(Cu, cps, sSu * sps);                     // cos(theta)
    double num = near ? S2 : Ca;
    double den = near ? C2 : Sa;
    double A = near ? flip_sign(M.A_near, sgn) : M.A_far;
    double B = near ? M.B_near : M.B_far;
    double inv = fast_rcp(den);
    v = (A * num) * inv;
    kk = (B * inv) * inv;
    if (LN) {
        ffar = near ? 1.0 : fabs(Sa);
        fnear = near ? fabs(C2) : 1.0;
        is_near = near ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// Shared-memory view of one CTA: G walkers, each with
//   tab [4 arrays][4 particles-in-block][nbp blocks]   (sa, ca, su, cu)
//   Q   [kmax+1 slots][4][nbp]   column partial sums of the drift
//   red [2][nbp]                 per-thread partials of E_L and ln|Psi|
// Particle p = 4 J + c is stored at [..][c][J]: the threads of a walker walk
// J, so every access is unit-stride across lanes (no bank conflicts).
// ---------------------------------------------------------------------------
struct GroupSmem {
    double *base;
    int nbp, kslots;
    __device__ __forceinline__ int walker_stride() const
    {
        return (16 + 4 * kslots + 2) * nbp;
    }
    __device__ __forceinline__ double *tab(int g, int a, int c) const
    {
        return base + g * walker_stride() + (a * 4 + c) * nbp;
    }
    __device__ __forceinline__ double *q(int g, int k, int c) const
    {
        return base + g * walker_stride() + (16 + k * 4 + c) * nbp;
    }
    __device__ __forceinline__ double *red(int g, int which) const
    {
        return base + g * walker_stride() + (16 + 4 * kslots + which) * nbp;
    }
};

__host__ __device__ inline int group_smem_doubles(int G, int nbp, int kslots)
{
    return G * (16 + 4 * kslots + 2) * nbp;
}

// Result of a walker-group evaluation, per thread.
struct EvalOut {
    double F[TB];       // drift of the thread's own particles
    double energy;      // local energy of the walker (valid on every thread)
    double lnpsi;       // ln|Psi| of the walker (valid on every thread)
};

// Evaluate drift, local energy (EF) and/or ln|Psi| (LN) of G walkers held by
// this CTA.  Thread (g, I) owns particles 4I..4I+3 of walker g, positions in
// z[] (entries >= nvalid are padding).  Every thread of the CTA must call
// this (it synchronises); `active` is false for surplus threads and for
// walkers that are not live.
template <bool LN, bool EF>
__device__ __forceinline__ void group_eval(const DevModel &M,
                                           const GroupSmem &sm, int g, int I,
                                           bool active, const double (&z)[TB],
                                           int nvalid, EvalOut &out)
{
    const int nb = M.nb, kmax = M.kmax;
    double rsa[TB], rca[TB], rsu[TB], rcu[TB];
    double F[TB];
    double e1 = 0.0, ln1 = 0.0;         // one-body partials
#pragma unroll
    for (int c = 0; c < TB; ++c) {
        F[c] = 0.0;
        rsa[c] = 0.0; rca[c] = 1.0; rsu[c] = 0.0; rcu[c] = 1.0;
    }
    if (active) {
#pragma unroll
        for (int c = 0; c < TB; ++c) {
            if (c < nvalid) {
                if (!M.is_ideal)
                    particle_tables(M, z[c], rsa[c], rca[c], rsu[c], rcu[c]);
                if (!M.is_free) {
                    OneBody ob = one_body<LN>(M, z[c]);
                    F[c] = ob.ldz;
                    e1 += ob.kin + ob.pot;
                    if (LN) ln1 += ob.lnf;
                }
            }
            sm.tab(g, 0, c)[I] = rsa[c];
            sm.tab(g, 1, c)[I] = rca[c];
            sm.tab(g, 2, c)[I] = rsu[c];
            sm.tab(g, 3, c)[I] = rcu[c];
        }
    }
    __syncthreads();

    double kin2 = 0.0;                  // sum over this thread's pairs
    double pf = 1.0, pn = 1.0;          // ln|f2| products (far / near)
    int ef = 0, en = 0, nnear = 0;
    if (active && !M.is_ideal) {
        const bool even = (nb & 1) == 0;
        for (int k = 0; k <= kmax; ++k) {
            // antipodal block column of an even ring: only half the rows
            if (k > 0 && even && k == kmax && I >= kmax) break;
            int J = I + k;
            if (J >= nb) J -= nb;
            const int nvj = min(TB, M.nop - TB * J);
            const double *tsa = sm.tab(g, 0, 0) + J;
            const int nbp = sm.nbp;
#pragma unroll 1
            for (int c2 = 0; c2 < TB; ++c2) {
                double jsa = tsa[(0 * 4 + c2) * nbp];
                double jca = tsa[(1 * 4 + c2) * nbp];
                double jsu = tsa[(2 * 4 + c2) * nbp];
                double jcu = tsa[(3 * 4 + c2) * nbp];
                double fc = 0.0;
#pragma unroll
                for (int c1 = 0; c1 < TB; ++c1) {
                    // diagonal block: c1 < c2 only; padding never counts
                    bool ok = (c1 < nvalid) && (c2 < nvj)
                              && (k > 0 || c1 < c2);
                    double v, kk, ffar, fnear;
                    int isn;
                    pair_term<LN>(M, rsa[c1], rca[c1], rsu[c1], rcu[c1],
                                  jsa, jca, jsu, jcu, v, kk, ffar, fnear,
                                  isn);
                    v = ok ? v : 0.0;
                    kk = ok ? kk : 0.0;
                    if (EF) {
                        F[c1] += v;
                        fc -= v;
                        kin2 += kk;
                    }
                    if (LN) {
                        pf *= ok ? ffar : 1.0;
                        pn *= ok ? fnear : 1.0;
                        nnear += ok ? isn : 0;
                    }
                }
                if (EF) sm.q(g, k, c2)[J] = fc;
                if (LN) { renorm(pf, ef); renorm(pn, en); }
            }
        }
    }
    __syncthreads();

    double epart = 0.0, lpart = 0.0;
    if (active) {
        if (EF && !M.is_ideal) {
            const bool even = (nb & 1) == 0;
            for (int k = 0; k <= kmax; ++k) {
                // slot k, column I was written by row block I - k
                if (k > 0 && even && k == kmax && I < kmax) continue;
#pragma unroll
                for (int c = 0; c < TB; ++c) F[c] += sm.q(g, k, c)[I];
            }
        }
        if (EF) {
            double f2 = 0.0;
#pragma unroll
            for (int c = 0; c < TB; ++c)
                if (c < nvalid) f2 = fma(F[c], F[c], f2);
            epart = e1 + 2.0 * kin2 - f2;
            sm.red(g, 0)[I] = epart;
        }
        if (LN) {
            lpart = ln1;
            if (!M.is_ideal)
                lpart += M.beta * (log(pf) + ef * LN2)
                         + (log(pn) + en * LN2) + nnear * M.ln_am;
            sm.red(g, 1)[I] = lpart;
        }
    }
    __syncthreads();
    double esum = 0.0, lsum = 0.0;
    if (active) {
        for (int t = 0; t < nb; ++t) {
            if (EF) esum += sm.red(g, 0)[t];
            if (LN) lsum += sm.red(g, 1)[t];
        }
    }
#pragma unroll
    for (int c = 0; c < TB; ++c) out.F[c] = F[c];
    out.energy = esum;
    out.lnpsi = lsum;
}

}  // namespace qmcb
