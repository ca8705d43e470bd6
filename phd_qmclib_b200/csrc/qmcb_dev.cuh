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
// from angle-addition identities, one reciprocal and no transcendental.  For
// the pair (i, j), d = z_i - z_j, r = |d_min|, sigma = sgn(d_min):
//   far  (r >= r_m):  sigma f2'/f2 = (pi/L) beta  cot(a_i - a_j)
//                     -f2''/f2 + (f2'/f2)^2 = (pi/L)^2 beta / sin^2(a_i - a_j)
//   near (r <  r_m):  sigma f2'/f2 = -k2 tan(u_i - u_j')
//                     -f2''/f2 + (f2'/f2)^2 = k2^2 / cos^2(u_i - u_j')
//     with u_j' = u_j + sigma psi_w, psi_0 = k2 r_off (|d| <= L/2),
//     psi_1 = k2 r_off - k2 L (wrapped pair): the minimum-image fold and the
//     sign of d become one of FOUR precomputed rotations of the column
//     particle's (sin u_j, cos u_j), picked by the sign bits of
//     sin(a_i - a_j) and cos(a_i - a_j).
// cot and 1/sin^2 have period L in d, so the far branch needs no fold at all,
// and r < r_m <=> |sin(a_i - a_j)| < sin(pi r_m / L) because r <= L/2.
// Both branches are brought to the same units (the column tables of the far
// branch are pre-scaled) so that one reciprocal yields
//   t = num/den  (drift contribution / -k2)   and   1/den^2  (kinetic / k2^2)
// with no per-pair constants.  Results agree with the reference to ~1e-15
// relative (tests/), far inside the 1e-12 budget of the north star.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmcb {

#ifndef QMCB_TB
#define QMCB_TB 4
#endif
constexpr int TB = QMCB_TB;           // particles owned by one thread
constexpr double LN2 = 0.693147180559945309417232121458;

// Node tables of the per-particle transcendentals (built once per model on
// the host, read through L1).  Every sin/cos/sinh/cosh a particle needs is a
// table entry at the nearest node rotated by the small remaining angle
// (|eps| <= EPS_MAX, short Taylor polynomials): ~12 fp64 instructions per
// angle instead of a ~100-instruction sincospi / exp.  Pure functions of z,
// so nothing is carried between time steps.
struct TrigTab {
    const double4 *zt;  // [nz + 1] node z_k = k L / nz:
                        //   (sin, cos)(pi z_k / L), (sin, cos)(k2 z_k)
    const double4 *ct;  // [nc + 1] node zc_j = j / nc of the unit cell:
                        //   (sin, cos)(k1 (zc_j - za/2)),
                        //   (sinh, cosh)(kp1 (zc_j - 1 + zb/2))
    double z_scale, eps_a, eps_u;       // nz / L, pi / nz, k2 L / nz
    double c_scale, eps_w, eps_b;       // nc, k1 / nc, kp1 / nc
    int nz, nc;
    // 32-bit word -> node index on the circle of 2 nz nodes
    __host__ __device__ double z_scale_circle() const
    {
        return 2.0 * nz / 4294967296.0;
    }
};
constexpr double TRIG_EPS_MAX = 0.0125;

struct DevModel {
    int nop;            // N
    int nb;             // ceil(N / 4): particle blocks == threads per walker
    int kmax;           // circulant half-width: nb / 2
    int is_free, is_ideal;
    int defects_sep;
    double L, inv_L;
    // one-body (Kronig-Penney cell: well [0, za], barrier (za, 1))
    double za, zb, k1, kp1, e0, v0, vdef, ln_cf;
    double k1_over_pi;
    // two-body
    double beta, ln_am, k2;
    double k2_over_pi;
    double inv_gam;                   // 1/gamma_f, gamma_f = (pi/L) sqrt(beta)/k2
    double mu;                        // -(pi/L) beta / k2
    double s_m_scaled;                // sin(pi r_m / L) / gamma_f
    double ln_gam;                    // ln gamma_f
    double cpsi[2], spsi[2];          // cos/sin(psi_w)
    double drift_unit;                // -k2       (1 if is_ideal)
    double kin_unit;                  // k2^2      (1 if is_ideal)
    double inv_drift_unit, half_inv_kin_unit;
    TrigTab tt;                       // zt == nullptr: no tables (exact path)
};

// Launch geometry shared by every walker-group kernel.
struct GroupGeom {
    int nthreads;       // CTA size (multiple of 32)
    int G;              // walkers per CTA
    int nbp;            // padded row length of the shared tables (>= nb, even)
    int kc;             // column-sum slots kept in shared memory at a time
    int tab_stride;     // doubles between the table regions of two walkers
    int q_stride;       // doubles between their column-sum regions
    int interleave;     // thread t -> (walker t % G, block t / G)
    int smem_bytes;
};

// Bank layout.  A warp reads element J = I + k (mod nb) of walker g for each of
// its lanes; 8 consecutive lanes form one 128-byte wavefront of a 16-byte
// access (16 lanes for 8-byte accesses), so consecutive lanes should land in
// consecutive 16-byte (8-byte) slots modulo 128 bytes.
//  * contiguous mapping (t = nb g + I): spacing the walker regions by
//    (nb elements mod 128 B) lets the lanes of the next walker continue the
//    sequence; the wrap of J inside a walker still costs one extra wavefront.
//  * interleaved mapping (t = G I + g, G odd): with the walker regions spaced
//    by s elements, G s = 1 (mod 8 resp. 16), slot(t) = G' t (mod 8/16) with
//    G' odd: conflict-free, and the wrap of J falls in one wavefront per CTA.
__host__ __device__ inline int bank_stride(int min_doubles, int want,
                                           int modulo)
{
    int s = min_doubles;
    while (s % modulo != want) ++s;
    return s;
}

__host__ __device__ inline int mod_inverse(int a, int m)
{
    for (int x = 1; x < m; ++x)
        if ((a * x) % m == 1) return x;
    return 0;
}

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
#ifdef QMCB_RCP_NEWTON2
    return fma(y, e, y);
#else
    double p = fma(e, e, e);
    return fma(y, p, y);
#endif
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
// A move leaves the interval by less than its length in all but pathological
// cases: those take two selects; anything else (and NaN) goes through the
// out-of-line fmod.
__device__ __noinline__ double recast_far(double x, double size)
{
    double r = fmod(x, size);
    if (r < 0.0) r += size;
    return r;
}

__device__ __forceinline__ double recast(double z, double z_min, double size)
{
    const double x = z - z_min;
    double r = x;                           // x in [0, size): exact
    r = (x < 0.0) ? x + size : r;           // Python: fmod(x) = x, then += b
    r = (x >= size) ? x - size : r;         // exact
    if (!(x >= -size && x < 2.0 * size)) r = recast_far(x, size);
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

// Four 32-bit draws of one Philox call.  Convention (shared with the oracle):
//   uniform  u = (x + 1/2) 2^-32                               in (0, 1)
//   normals  (Box-Muller) r_a = sqrt(-2 ln((x0 + 1/2) 2^-32)),
//            (n0, n1) = r_a (cos, sin)(2 pi x1 2^-32); (n2, n3) from x2, x3.
__device__ __forceinline__ void rng_words4(uint64_t seed, uint32_t c0,
                                           uint32_t c1, uint32_t c2,
                                           uint32_t stream, uint32_t (&x)[4])
{
    x[0] = c0; x[1] = c1; x[2] = c2; x[3] = stream;
    philox4x32_10(x, (uint32_t) seed, (uint32_t) (seed >> 32));
}

__device__ __forceinline__ double u32_open(uint32_t x)
{
    return ((double) x + 0.5) * (1.0 / 4294967296.0);
}

__device__ __forceinline__ void rng_uniform4(uint64_t seed, uint32_t c0,
                                             uint32_t c1, uint32_t c2,
                                             uint32_t stream, double (&u)[4])
{
    uint32_t x[4];
    rng_words4(seed, c0, c1, c2, stream, x);
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = u32_open(x[i]);
}

// one 32-byte table entry through the read-only path (two LDG.128)
__device__ __forceinline__ double4 ldg4(const double4 *p)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// sin(e) and cos(e) - 1 for |e| <= TRIG_EPS_MAX (truncation < 1e-17)
__device__ __forceinline__ void small_sincos(double e, double &se, double &dc)
{
    const double s2 = e * e;
    const double p = fma(s2, 1.0 / 120.0, -1.0 / 6.0);
    se = fma(e * s2, p, e);
    double q = fma(s2, -1.0 / 720.0, 1.0 / 24.0);
    q = fma(s2, q, -0.5);
    dc = s2 * q;
}

// sinh(e) and cosh(e) - 1, same range
__device__ __forceinline__ void small_sinhcosh(double e, double &se,
                                               double &dc)
{
    const double s2 = e * e;
    const double p = fma(s2, 1.0 / 120.0, 1.0 / 6.0);
    se = fma(e * s2, p, e);
    double q = fma(s2, 1.0 / 720.0, 1.0 / 24.0);
    q = fma(s2, q, 0.5);
    dc = s2 * q;
}

// (sin, cos)(t0 + e) from (s0, c0) = (sin, cos)(t0)
__device__ __forceinline__ void rotate_by(double s0, double c0, double se,
                                          double dc, double &s, double &c)
{
    s = fma(c0, se, fma(s0, dc, s0));
    c = fma(-s0, se, fma(c0, dc, c0));
}

// nearest integer of x (|x| < 2^31) and x minus it, without F2I / I2F
__device__ __forceinline__ int nearest_node(double x, double &frac)
{
    const double magic = 6755399441055744.0;        // 1.5 * 2^52
    const double m = x + magic;
    frac = x - (m - magic);
    return __double2loint(m);
}

template <bool FAST>
__device__ __forceinline__ void rng_normal4(const TrigTab &tt, uint64_t seed,
                                            uint32_t c0, uint32_t c1,
                                            uint32_t c2, uint32_t stream,
                                            double (&n)[4])
{
    uint32_t x[4];
    rng_words4(seed, c0, c1, c2, stream, x);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const double r = sqrt(-2.0 * log(u32_open(x[2 * h])));
        double s, c;
        if (FAST) {
            // angle 2 pi v = pi (2 v): node k' of 2 nz nodes on the circle;
            // the table holds the upper half, the lower half is its negative
            double fr;
            int k = nearest_node((double) x[2 * h + 1]
                                 * (tt.z_scale_circle()), fr);
            const bool lower = k >= tt.nz;
            k -= lower ? tt.nz : 0;     // k' = 2 nz -> -(sin, cos)(pi) = (0, 1)
            const double4 e = ldg4(tt.zt + k);
            double se, dc;
            small_sincos(fr * tt.eps_a, se, dc);
            rotate_by(e.x, e.y, se, dc, s, c);
            s = lower ? -s : s;
            c = lower ? -c : c;
        } else {
            sincospi((double) x[2 * h + 1] * (2.0 / 4294967296.0), &s, &c);
        }
        n[2 * h] = r * c;
        n[2 * h + 1] = r * s;
    }
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
    const bool barrier = M.za < zc;
    if (!LN) {
        // Branch-free: both regions are evaluated and one is selected, so
        // that the four particles of a thread interleave instead of taking
        // divergent branches one after the other.  tanh from one exp and one
        // reciprocal: absolute error ~1e-16, which is what matters for a
        // term added to O(1) drifts and energies.
        double argb = M.kp1 * (zc - 1.0 + 0.5 * M.zb);
        double ex = exp(-2.0 * fabs(argb));
        double th = copysign((1.0 - ex) * fast_rcp(1.0 + ex), argb);
        double ldz_b = M.kp1 * th;
        double s, c;
        sincospi(M.k1_over_pi * (zc - 0.5 * M.za), &s, &c);
        double ldz_w = -M.k1 * (s * fast_rcp(c));
        o.ldz = barrier ? ldz_b : ldz_w;
        o.kin = fma(o.ldz, o.ldz, barrier ? -(M.v0 - M.e0) : M.e0);
        // mrbp_qmc/model.py:533-551: every defects_sep-th cell is a defect
        double vb = M.vdef;
        if (M.defects_sep != 1
            && ((int) n_cell % M.defects_sep) != 0)
            vb = M.v0;
        o.pot = barrier ? vb : 0.0;
        o.lnf = 0.0;
        return o;
    }
    if (barrier) {
        double arg = M.kp1 * (zc - 1.0 + 0.5 * M.zb);
        double ex = exp(-2.0 * fabs(arg));
        double th = copysign((1.0 - ex) * fast_rcp(1.0 + ex), arg);
        o.ldz = M.kp1 * th;
        o.kin = -(M.v0 - M.e0) + o.ldz * o.ldz;
        bool defect = (M.defects_sep == 1)
                      || (((int) n_cell % M.defects_sep) == 0);
        o.pot = defect ? M.vdef : M.v0;
        o.lnf = log(cosh(arg));
    } else {                                        // well
        double s, c;
        sincospi(M.k1_over_pi * (zc - 0.5 * M.za), &s, &c);
        o.ldz = -M.k1 * (s * fast_rcp(c));
        o.kin = M.e0 + o.ldz * o.ldz;
        o.pot = 0.0;
        o.lnf = M.ln_cf + log(fabs(c));
    }
    return o;
}

// Table form of one_body<false> (no ln f1): the well angle and the barrier
// argument are both continued over the whole unit cell, so one node entry
// serves either region and the region is a select, not a branch.
__device__ __forceinline__ OneBody one_body_fast(const DevModel &M, double z)
{
    const TrigTab &tt = M.tt;
    OneBody o;
    const double n_cell = floor(z);
    const double zc = z - n_cell;                   // z mod 1, exact
    const bool barrier = M.za < zc;
    double fr;
    int j = nearest_node(zc * tt.c_scale, fr);
    j = min(max(j, 0), tt.nc);
    const double4 e = ldg4(tt.ct + j);
    double se, dc, sw, cw;
    small_sincos(fr * tt.eps_w, se, dc);
    rotate_by(e.x, e.y, se, dc, sw, cw);
    small_sinhcosh(fr * tt.eps_b, se, dc);
    const double sh = fma(e.w, se, fma(e.z, dc, e.z));
    const double ch = fma(e.z, se, fma(e.w, dc, e.w));
    // f1'/f1 = kp1 tanh(.) (barrier) | -k1 tan(.) (well): one reciprocal
    const double num = barrier ? M.kp1 * sh : -M.k1 * sw;
    const double den = barrier ? ch : cw;
    o.ldz = num * fast_rcp(den);
    // mrbp_qmc/model.py:533-551: every defects_sep-th cell is a defect
    double vb = M.vdef;
    if (M.defects_sep != 1 && ((int) n_cell % M.defects_sep) != 0)
        vb = M.v0;
    o.kin = fma(o.ldz, o.ldz, barrier ? -(M.v0 - M.e0) : M.e0);
    o.pot = barrier ? vb : 0.0;
    o.lnf = 0.0;
    return o;
}

// Table form of one_body<true>: |f1| comes back as a factor (cosh in the
// barrier, cos in the well, where ln cf is added per well particle by the
// caller) so that a thread takes ONE logarithm for its four particles.
__device__ __forceinline__ OneBody one_body_fast_ln(const DevModel &M,
                                                    double z, double &f1abs,
                                                    int &in_well)
{
    const TrigTab &tt = M.tt;
    OneBody o;
    const double n_cell = floor(z);
    const double zc = z - n_cell;
    const bool barrier = M.za < zc;
    double fr;
    int j = nearest_node(zc * tt.c_scale, fr);
    j = min(max(j, 0), tt.nc);
    const double4 e = ldg4(tt.ct + j);
    double se, dc, sw, cw;
    small_sincos(fr * tt.eps_w, se, dc);
    rotate_by(e.x, e.y, se, dc, sw, cw);
    small_sinhcosh(fr * tt.eps_b, se, dc);
    const double sh = fma(e.w, se, fma(e.z, dc, e.z));
    const double ch = fma(e.z, se, fma(e.w, dc, e.w));
    const double num = barrier ? M.kp1 * sh : -M.k1 * sw;
    const double den = barrier ? ch : cw;
    o.ldz = num * fast_rcp(den);
    double vb = M.vdef;
    if (M.defects_sep != 1 && ((int) n_cell % M.defects_sep) != 0)
        vb = M.v0;
    o.kin = fma(o.ldz, o.ldz, barrier ? -(M.v0 - M.e0) : M.e0);
    o.pot = barrier ? vb : 0.0;
    o.lnf = 0.0;
    f1abs = fabs(den);
    in_well = barrier ? 0 : 1;
    return o;
}

// sin/cos tables of one particle from the node table (z in [0, L])
__device__ __forceinline__ void particle_tables_fast(const DevModel &M,
                                                     double z, double &sa,
                                                     double &ca, double &su,
                                                     double &cu)
{
    const TrigTab &tt = M.tt;
    double fr;
    int k = nearest_node(z * tt.z_scale, fr);
    k = min(max(k, 0), tt.nz);
    const double4 e = ldg4(tt.zt + k);
    double se, dc;
    small_sincos(fr * tt.eps_a, se, dc);
    rotate_by(e.x, e.y, se, dc, sa, ca);
    small_sincos(fr * tt.eps_u, se, dc);
    rotate_by(e.z, e.w, se, dc, su, cu);
}

// sin/cos tables of one particle
__device__ __forceinline__ void particle_tables(const DevModel &M, double z,
                                                double &sa, double &ca,
                                                double &su, double &cu)
{
    sincospi(z * M.inv_L, &sa, &ca);
    sincospi(z * M.k2_over_pi, &su, &cu);
}

// ---------------------------------------------------------------------------
// Shared-memory view of one CTA: G walkers, each with (nbp doubles per row)
//   A1  [4][nbp] double2   (sin a_j, cos a_j) / gamma_f   far branch
//   V   [4 variants][4][nbp] double2  (sin, cos)(u_j + sigma psi_w)
//   Q   [kc][4][nbp]       column partial sums of the drift; its first two
//                          rows double as the per-thread partials of E_L and
//                          ln|Psi| in the final reduction
// Particle p = 4 J + c is stored at [..][c][J]: the threads of a walker walk
// J, so accesses are unit-stride across lanes.  nbp is even, which makes the
// variant stride a multiple of 128 bytes: lanes that pick different variants
// still hit distinct banks.
// ---------------------------------------------------------------------------
// Rows per walker: far table 8 + variants 32; the final reductions reuse the
// column-sum rows.  (A second, pre-scaled copy of the far table would save
// one multiply per pair but costs more in shared-memory traffic and in
// column-sum slots than it gains: measured -4 %.)
constexpr int TAB_ROWS = 10 * TB;
constexpr int RED_ROWS = 0;

struct GroupSmem {
    double *base;
    int nbp, kc, G, tab_stride, q_stride;
    __device__ __forceinline__ double *tab(int g) const
    {
        return base + g * tab_stride;
    }
    __device__ __forceinline__ double *qreg(int g) const
    {
        return base + G * tab_stride + g * q_stride;
    }
    __device__ __forceinline__ double2 *a1(int g, int c) const
    {
        return reinterpret_cast<double2 *>(tab(g)) + c * nbp;
    }
    __device__ __forceinline__ double2 *var(int g, int v, int c) const
    {
        return reinterpret_cast<double2 *>(tab(g) + 2 * TB * nbp)
               + (v * TB + c) * nbp;
    }
    __device__ __forceinline__ double *q(int g, int k, int c) const
    {
        return qreg(g) + (k * TB + c) * nbp;
    }
    __device__ __forceinline__ double *red(int g, int which) const
    {
        return qreg(g) + which * nbp;       // after the last column-sum fold
    }
};

__host__ __device__ inline int group_tab_stride(int nbp, int nb, int G,
                                                bool interleave)
{
    // 16-byte elements: slot = doubles / 2, modulo 8 slots
    int want = interleave ? mod_inverse(G % 8, 8) : nb % 8;
    return bank_stride(TAB_ROWS * nbp + ((TAB_ROWS * nbp) & 1), 2 * want, 16);
}

__host__ __device__ inline int group_q_stride(int nbp, int nb, int kc, int G,
                                              bool interleave)
{
    // 8-byte elements, modulo 16 slots
    int want = interleave ? mod_inverse(G % 16, 16) : nb % 16;
    return bank_stride((TB * kc + RED_ROWS) * nbp, want, 16);
}

__host__ __device__ inline int group_smem_doubles(int G, int nbp, int nb,
                                                  int kc, bool interleave)
{
    return G * (group_tab_stride(nbp, nb, G, interleave)
                + group_q_stride(nbp, nb, kc, G, interleave));
}

// Result of a walker-group evaluation, per thread.
struct EvalOut {
    double F[TB];       // drift of the thread's own particles
    double energy;      // local energy of the walker (on every thread of the
    double lnpsi;       // walker if BCAST, else on its thread I == 0) / ln|Psi|
};

// Accumulators of one thread over its pair tiles.
struct PairAcc {
    double T[TB];       // sum_j t_ij of the thread's rows
    double K;           // sum over the thread's pairs of 1/den^2
    double pf, pn;      // products of |den| over far / near pairs
    int ef, en;         // their binary exponents
    int nnear, npair;
};

// One 4x4 tile: rows = the thread's particles (tables in registers), columns
// = block J of the same walker (tables in shared memory).  MASK: the tile is
// the diagonal block (pairs c1 < c2 only) or touches padding particles.
template <bool LN, bool EF, bool MASK>
__device__ __forceinline__ void pair_tile(
    const DevModel &M, const GroupSmem &sm, int g, int J, int qslot,
    bool diag, int nvalid, int nvj, const double (&rsa)[TB],
    const double (&rca)[TB], const double (&rsu)[TB],
    const double (&rcu)[TB], PairAcc &acc)
{
    const int nbp = sm.nbp;
    const int cstride = nbp * (int) sizeof(double2);    // next column particle
    const unsigned vstride = (unsigned) TB * (unsigned) cstride;   // next variant
    const char *pa1 = reinterpret_cast<const char *>(sm.a1(g, 0) + J);
    const char *pv = reinterpret_cast<const char *>(sm.var(g, 0, 0) + J);
    double *pq = sm.q(g, qslot, 0) + J;
    const double s_m = M.s_m_scaled;
    double2 A1 = *reinterpret_cast<const double2 *>(pa1);
    const double mu = M.mu;
#pragma unroll 1
    for (int c2 = 0; c2 < TB; ++c2) {
        // phase 1: far branch in near units for the four rows,
        //   den_f = sin(a_i - a_j) / gamma_f,
        //   num_f = (mu_f / gamma_f) cos(a_i - a_j),  mu_f < 0,
        // and the column-table variant each pair needs: bit 1 = unwrapped
        // (cos > 0 <=> num_f < 0), bit 0 = sign bit of den_f
        double den_f[TB], num_f[TB];
        double2 V[TB];
#pragma unroll
        for (int c1 = 0; c1 < TB; ++c1) {
            den_f[c1] = fma(rsa[c1], A1.y, -(rca[c1] * A1.x));
            num_f[c1] = mu * fma(rca[c1], A1.y, rsa[c1] * A1.x);
        }
#pragma unroll
        for (int c1 = 0; c1 < TB; ++c1) {
            unsigned hn = (unsigned) __double2hiint(num_f[c1]);
            unsigned hd = (unsigned) __double2hiint(den_f[c1]);
            unsigned v = ((hn >> 31) << 1) + (hd >> 31);
            V[c1] = *reinterpret_cast<const double2 *>(pv + v * vstride);
        }
        // next column particle's far tables, in flight during phase 2
        if (c2 + 1 < TB) {
            pa1 += cstride;
            A1 = *reinterpret_cast<const double2 *>(pa1);
        }
        pv += cstride;
        // phase 2: near branch, select, one reciprocal per pair
        double fc = 0.0;
#pragma unroll
        for (int c1 = 0; c1 < TB; ++c1) {
            bool near = fabs(den_f[c1]) < s_m;
            double num_n = fma(rsu[c1], V[c1].y, -(rcu[c1] * V[c1].x));
            double den_n = fma(rcu[c1], V[c1].y, rsu[c1] * V[c1].x);
            double num = near ? num_n : num_f[c1];
            double den = near ? den_n : den_f[c1];
            double inv = fast_rcp(den);
            double t = num * inv;
            if (MASK) {
                bool ok = (c1 < nvalid) && (c2 < nvj) && (!diag || c1 < c2);
                t = ok ? t : 0.0;
                inv = ok ? inv : 0.0;
                if (LN) {
                    den_f[c1] = ok ? den_f[c1] : 1.0;
                    den = ok ? den : 1.0;
                    acc.npair += ok ? 1 : 0;
                    acc.nnear += (ok && near) ? 1 : 0;
                }
            } else if (LN) {
                acc.npair += 1;
                acc.nnear += near ? 1 : 0;
            }
            if (EF) {
                acc.T[c1] += t;
                fc -= t;
                acc.K = fma(inv, inv, acc.K);
            }
            if (LN) {
                acc.pf *= near ? 1.0 : fabs(den_f[c1]);
                acc.pn *= fabs(den);       // every pair (see group_eval)
            }
        }
        if (EF) { *pq = fc; pq += nbp; }
        if (LN) { renorm(acc.pf, acc.ef); renorm(acc.pn, acc.en); }
    }
}

// The diagonal tile of a full block: the six pairs c1 < c2 among the thread's
// own four particles, unrolled, without masks; both ends of a pair are rows of
// this thread, so the "column" contribution goes straight into T (no
// shared-memory slot, nothing to fold).  The generic masked tile spends 16
// pair slots on these 6 pairs.
template <bool RAGGED>
__device__ __forceinline__ void pair_diag(
    const DevModel &M, const GroupSmem &sm, int g, int I, int nvalid,
    const double (&rsa)[TB], const double (&rca)[TB],
    const double (&rsu)[TB], const double (&rcu)[TB], PairAcc &acc)
{
    const unsigned vstride = (unsigned) TB * (unsigned) sm.nbp * (unsigned) sizeof(double2);
    const double s_m = M.s_m_scaled, mu = M.mu;
#pragma unroll
    for (int c2 = 1; c2 < TB; ++c2) {
        if (RAGGED && c2 >= nvalid) break;      // last, partly filled block
        const double2 A1 = sm.a1(g, c2)[I];
        const char *pv = reinterpret_cast<const char *>(sm.var(g, 0, c2) + I);
#pragma unroll
        for (int c1 = 0; c1 < c2; ++c1) {
            const double den_f = fma(rsa[c1], A1.y, -(rca[c1] * A1.x));
            const double num_f = mu * fma(rca[c1], A1.y, rsa[c1] * A1.x);
            const unsigned hn = (unsigned) __double2hiint(num_f);
            const unsigned hd = (unsigned) __double2hiint(den_f);
            const unsigned v = ((hn >> 31) << 1) + (hd >> 31);
            const double2 V =
                *reinterpret_cast<const double2 *>(pv + v * vstride);
            const bool near = fabs(den_f) < s_m;
            const double num_n = fma(rsu[c1], V.y, -(rcu[c1] * V.x));
            const double den_n = fma(rcu[c1], V.y, rsu[c1] * V.x);
            const double num = near ? num_n : num_f;
            const double den = near ? den_n : den_f;
            const double inv = fast_rcp(den);
            const double t = num * inv;
            acc.T[c1] += t;
            acc.T[c2] -= t;
            acc.K = fma(inv, inv, acc.K);
        }
    }
}

// Off-diagonal 4x4 tile of the drift / energy kernels (no ln|Psi|), with the
// ragged last block (N not a multiple of 4) at full speed:
//  * a ragged COLUMN block just ends the column loop early (`ncol` < 4): the
//    lanes of a warp that meet another block idle for those iterations, no
//    extra instruction is issued;
//  * a ragged ROW block (the one thread per walker that owns it) nulls the
//    reciprocal of its padding rows (two FSEL per pair; the numerators are
//    finite, so t = num * 0 and 1/den^2 vanish).  ROWMASK is chosen per WARP
//    (any lane owns a ragged block), so that the lanes of a warp never run
//    two different tile routines one after the other -- that serialisation
//    made the warp of the ragged block the straggler of every CTA barrier
//    (N = 50: 1.09 ms per step against 0.83 ms for N = 52).
// FULL: the model has no ragged block at all (N a multiple of 4, a uniform
// condition): compile-time column count.
template <bool LN, bool EF, bool ROWMASK, bool FULL = false>
__device__ __forceinline__ void pair_tile_lean(
    const DevModel &M, const GroupSmem &sm, int g, int J, int qslot,
    int nvalid, int ncol, const double (&rsa)[TB], const double (&rca)[TB],
    const double (&rsu)[TB], const double (&rcu)[TB], PairAcc &acc)
{
    const int nbp = sm.nbp;
    const int cstride = nbp * (int) sizeof(double2);    // next column particle
    const unsigned vstride = (unsigned) TB * (unsigned) cstride;   // next variant
    const char *pa1 = reinterpret_cast<const char *>(sm.a1(g, 0) + J);
    const char *pv = reinterpret_cast<const char *>(sm.var(g, 0, 0) + J);
    double *pq = sm.q(g, qslot, 0) + J;
    const double s_m = M.s_m_scaled;
    double2 A1 = *reinterpret_cast<const double2 *>(pa1);
    const double mu = M.mu;
#pragma unroll 1
    for (int c2 = 0; c2 < (FULL ? TB : ncol); ++c2) {
        double den_f[TB], num_f[TB];
        double2 V[TB];
#pragma unroll
        for (int c1 = 0; c1 < TB; ++c1) {
            den_f[c1] = fma(rsa[c1], A1.y, -(rca[c1] * A1.x));
            num_f[c1] = mu * fma(rca[c1], A1.y, rsa[c1] * A1.x);
        }
#pragma unroll
        for (int c1 = 0; c1 < TB; ++c1) {
            unsigned hn = (unsigned) __double2hiint(num_f[c1]);
            unsigned hd = (unsigned) __double2hiint(den_f[c1]);
            unsigned v = ((hn >> 31) << 1) + (hd >> 31);
            V[c1] = *reinterpret_cast<const double2 *>(pv + v * vstride);
        }
        if (c2 + 1 < (FULL ? TB : ncol)) {
            pa1 += cstride;
            A1 = *reinterpret_cast<const double2 *>(pa1);
        }
        pv += cstride;
        double fc = 0.0;
#pragma unroll
        for (int c1 = 0; c1 < TB; ++c1) {
            bool near = fabs(den_f[c1]) < s_m;
            double num_n = fma(rsu[c1], V[c1].y, -(rcu[c1] * V[c1].x));
            double den_n = fma(rcu[c1], V[c1].y, rsu[c1] * V[c1].x);
            double num = near ? num_n : num_f[c1];
            double den = near ? den_n : den_f[c1];
            double inv = fast_rcp(den);
            if (ROWMASK && c1 > 0) {
                const bool ok = c1 < nvalid;
                inv = ok ? inv : 0.0;
                if (LN) {
                    den_f[c1] = ok ? den_f[c1] : 1.0;
                    den = ok ? den : 1.0;
                    acc.nnear += (ok && near) ? 1 : 0;
                }
            } else if (LN) {
                acc.nnear += near ? 1 : 0;
            }
            if (EF) {
                double t = num * inv;
                acc.T[c1] += t;
                fc -= t;
                acc.K = fma(inv, inv, acc.K);
            }
            if (LN) {
                // pn collects EVERY pair's denominator here (the selected
                // one needs no second select); the far ones are divided out
                // in the logarithm (group_eval)
                acc.pf *= near ? 1.0 : fabs(den_f[c1]);
                acc.pn *= fabs(den);
            }
        }
        if (EF) { *pq = fc; pq += nbp; }
        if (LN) { renorm(acc.pf, acc.ef); renorm(acc.pn, acc.en); }
    }
    if (LN) acc.npair += FULL ? TB * TB : nvalid * ncol;
}

// Evaluate drift, local energy (EF) and/or ln|Psi| (LN) of G walkers held by
// this CTA.  Thread (g, I) owns particles 4I..4I+3 of walker g, positions in
// z[] (entries >= nvalid are padding).  Every thread of the CTA must call
// this (it synchronises); `active` is false for surplus threads and for
// walkers that are not live.
// BCAST: E_L and ln|Psi| of the walker are needed on every one of its threads
// (the Metropolis test of the VMC kernel); otherwise only thread I == 0 adds
// up the per-thread partials.
template <bool LN, bool EF, bool BCAST = true, bool FAST = false>
__device__ __forceinline__ void group_eval(const DevModel &M,
                                           const GroupSmem &sm, int g, int I,
                                           bool active, const double (&z)[TB],
                                           int nvalid, EvalOut &out)
{
    const int nb = M.nb, kmax = M.kmax, kc = sm.kc;
    double rsa[TB], rca[TB], rsu[TB], rcu[TB];
    PairAcc acc;
    // The one-body terms ride in the pair accumulators (in their units):
    // T starts at f1'/f1 / drift_unit, K at (kin1 + V) / (2 kin_unit).
    acc.K = 0.0; acc.pf = 1.0; acc.pn = 1.0;
    acc.ef = 0; acc.en = 0; acc.nnear = 0; acc.npair = 0;
    double ln1 = 0.0;
#pragma unroll
    for (int c = 0; c < TB; ++c) {
        acc.T[c] = 0.0;
        rsa[c] = 0.0; rca[c] = 1.0; rsu[c] = 0.0; rcu[c] = 1.0;
    }
    if (active) {
        double e1 = 0.0, p1 = 1.0;      // p1: product of |f1| (FAST && LN)
        int nwell = 0;
#pragma unroll
        for (int c = 0; c < TB; ++c) {
            if (c < nvalid) {
                if (!M.is_ideal) {
                    if (FAST)
                        particle_tables_fast(M, z[c], rsa[c], rca[c], rsu[c],
                                             rcu[c]);
                    else
                        particle_tables(M, z[c], rsa[c], rca[c], rsu[c],
                                        rcu[c]);
                }
                if (!M.is_free) {
                    OneBody ob;
                    if (FAST && LN) {
                        double f1;
                        int w;
                        ob = one_body_fast_ln(M, z[c], f1, w);
                        p1 *= f1;
                        nwell += w;
                    } else if (FAST) {
                        ob = one_body_fast(M, z[c]);
                    } else {
                        ob = one_body<LN>(M, z[c]);
                    }
                    acc.T[c] = ob.ldz * M.inv_drift_unit;
                    e1 += ob.kin + ob.pot;
                    if (LN) ln1 += ob.lnf;
                }
            }
            if (!M.is_ideal) {
                sm.a1(g, c)[I] = make_double2(rsa[c] * M.inv_gam,
                                              rca[c] * M.inv_gam);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    // v = 2 * unwrapped + [sin(a_i - a_j) < 0]: the two sign
                    // bits as they come; sigma > 0 <=> the bits differ
                    // (unwrapped: sin > 0; wrapped: sin < 0);
                    // u_j' = u_j + sigma psi
                    const int w = (v & 2) ? 0 : 1;
                    const double cp = M.cpsi[w];
                    const double sp = (((v >> 1) ^ v) & 1) ? M.spsi[w]
                                                           : -M.spsi[w];
                    sm.var(g, v, c)[I] = make_double2(
                        fma(rsu[c], cp, rcu[c] * sp),
                        fma(rcu[c], cp, -(rsu[c] * sp)));
                }
            }
        }
        acc.K = e1 * M.half_inv_kin_unit;
        // four factors of at most cosh(kp1 zb / 2) each: no overflow
        // (the tables are not built for kp1 > 600)
        if (FAST && LN && !M.is_free) ln1 = log(p1) + nwell * M.ln_cf;
    }
    __syncthreads();

    double Tq[TB] = {};   // column sums received from others
    const bool pairs = active && !M.is_ideal;
    const bool even = (nb & 1) == 0;
    // drift / energy kernels: the diagonal tile adds both ends of its pairs
    // to the thread's own rows and needs no column-sum slot, so the slots
    // (and the barrier pairs around their fold) serve k = 1 .. kmax only
    constexpr bool diag_direct = !LN && EF;
    if (diag_direct && pairs) {
        if (nvalid == TB)
            pair_diag<false>(M, sm, g, I, nvalid, rsa, rca, rsu, rcu, acc);
        else
            pair_diag<true>(M, sm, g, I, nvalid, rsa, rca, rsu, rcu, acc);
    }
    // per-thread constants of the tile loop: the last tile this thread
    // computes / folds (antipodal block column of an even ring: the first
    // half of the row blocks computes it, the second half receives it) and
    // the ragged last block (N not a multiple of 4)
    const int k_last_row = (even && I >= kmax) ? kmax - 1 : kmax;
    const int k_skip_col = (even && I < kmax) ? kmax : -1;
    const int j_ragged = (M.nop % TB) ? nb - 1 : -1;
    const int n_ragged = M.nop - TB * (nb - 1);     // particles of that block
    const bool row_ragged = nvalid < TB;
    // one tile routine per warp (see pair_tile_ef); every thread of the CTA
    // is here, active or not
    const bool warp_rowmask = __any_sync(0xffffffffu, pairs && row_ragged);
    for (int k0 = diag_direct ? 1 : 0; k0 <= kmax; k0 += kc) {
        const int k1 = min(k0 + kc, kmax + 1);
        if (pairs) {
            const int ke = min(k1, k_last_row + 1);
            for (int k = k0; k < ke; ++k) {
                int J = I + k;
                if (J >= nb) J -= nb;
                const int ncol = (J == j_ragged) ? n_ragged : TB;
                if (k == 0) {
                    // the diagonal block of the ln|Psi| kernels: pairs
                    // c1 < c2 only, through the generic masked tile
                    pair_tile<LN, EF, true>(M, sm, g, J, k - k0, true, nvalid,
                                            ncol, rsa, rca, rsu, rcu, acc);
                } else if (j_ragged < 0) {
                    pair_tile_lean<LN, EF, false, true>(
                        M, sm, g, J, k - k0, TB, TB, rsa, rca, rsu, rcu, acc);
                } else if (warp_rowmask) {
                    pair_tile_lean<LN, EF, true>(M, sm, g, J, k - k0, nvalid,
                                                 ncol, rsa, rca, rsu, rcu,
                                                 acc);
                } else {
                    pair_tile_lean<LN, EF, false>(M, sm, g, J, k - k0, nvalid,
                                                  ncol, rsa, rca, rsu, rcu,
                                                  acc);
                }
            }
        }
        if (EF) {
            __syncthreads();
            if (pairs) {
                for (int k = k0; k < k1; ++k) {
                    // slot k, column I was written by row block I - k
                    if (k == k_skip_col) continue;
#pragma unroll
                    for (int c = 0; c < TB; ++c)
                        Tq[c] += sm.q(g, k - k0, c)[I];
                }
            }
            __syncthreads();    // the reductions below reuse these rows
        }
    }
    if (!EF) __syncthreads();

    double epart = 0.0, lpart = 0.0;
    double F[TB] = {};
    if (active) {
        if (EF) {
            double f2 = 0.0;
#pragma unroll
            for (int c = 0; c < TB; ++c) {
                F[c] = M.drift_unit * (acc.T[c] + Tq[c]);
                if (c < nvalid) f2 = fma(F[c], F[c], f2);
            }
            epart = 2.0 * M.kin_unit * acc.K - f2;
            sm.red(g, 0)[I] = epart;
        }
        if (LN) {
            lpart = ln1;
            if (!M.is_ideal) {
                const int nfar = acc.npair - acc.nnear;
                // pn holds the denominators of ALL pairs, pf those of the
                // far ones: ln(near product) = ln pn - ln pf
                const double lf = log(acc.pf) + acc.ef * LN2;
                lpart += M.beta * (lf + nfar * M.ln_gam)
                         + ((log(acc.pn) + acc.en * LN2) - lf)
                         + acc.nnear * M.ln_am;
            }
            sm.red(g, 1)[I] = lpart;
        }
    }
    __syncthreads();
    double esum = 0.0, lsum = 0.0;
    if (active && (BCAST || I == 0)) {
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
