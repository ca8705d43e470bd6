// Estimator kernels of the DMC path: static structure factor S(k) and
// one-body density, mixed or pure (forward walking).
//
// Reference (paths relative to src/phd_qmclib/ of PhD-QMCLib):
//   rho_k of one configuration        qmc_base/jastrow/model.py:968-1004
//   S(k) per step, forward walking    qmc_base/jastrow/dmc.py:363-573
//   momenta k_m = m 2 pi / L, m>=0    mrbp_qmc/dmc.py:595-641
//   density per step                  mrbp_qmc/dmc.py:472-547,
//                                     qmc_base/jastrow/dmc.py:195-302
// Both estimators look at the "actual" population of a step: slot s holds the
// pre-move configuration of its parent, confs[ref[s]].
#pragma once
#include "qmcb_dev.cuh"

namespace qmcb {

constexpr int SSF_CHUNK = 16;       // modes advanced by one thread
constexpr int SSF_THREADS = 128;

struct SsfArgs {
    const double *confs;    // [*][2][N]
    const int *ref;         // slot -> row of confs, or null (identity)
    const int *W_dev;       // live walkers (device scalar) or null
    long long W_host;       // used when W_dev == null
    int N, M;
    double two_over_L;
    double *out;            // [W][M][3]  (|rho|^2, Re, Im) (+ prev if pure)
    const double *prev;     // [cap][M][3] or null
    int accumulate;         // out = value + prev[ref[s]]  (pure, step < pfw)
};

// Thread (g, j): walker g of the CTA, modes [16 j, 16 j + 16).  For every
// particle the phase e^{i k_m z} is seeded exactly at the first mode of the
// chunk and advanced by a three-term recurrence in cos(2 pi z / L).
__global__ void __launch_bounds__(SSF_THREADS)
ssf_eval_kernel(SsfArgs a, int G, int nchunk)
{
    extern __shared__ __align__(16) double ssf_smem[];
    const long long W = a.W_dev ? (long long) *a.W_dev : a.W_host;
    const int N = a.N;
    double *sx = ssf_smem;                  // [G][N]  x = 2 z / L
    double *sc1 = sx + (size_t) G * N;      // [G][N]  cos(pi x)
    double *ss1 = sc1 + (size_t) G * N;     // [G][N]  sin(pi x)
    for (long long s0 = (long long) blockIdx.x * G; s0 < W;
         s0 += (long long) gridDim.x * G) {
        for (int e = threadIdx.x; e < G * N; e += blockDim.x) {
            int g = e / N, i = e - g * N;
            long long s = s0 + g;
            double x = 0.0, c = 1.0, sn = 0.0;
            if (s < W) {
                long long r = a.ref ? (long long) a.ref[s] : s;
                x = a.confs[r * 2 * N + i] * a.two_over_L;
                sincospi(x, &sn, &c);
            }
            sx[e] = x; sc1[e] = c; ss1[e] = sn;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < G * nchunk; t += blockDim.x) {
            const int g = t / nchunk, j = t - g * nchunk;
            const long long s = s0 + g;
            if (s >= W) continue;
            const int m0 = j * SSF_CHUNK;
            double re[SSF_CHUNK], im[SSF_CHUNK];
#pragma unroll
            for (int q = 0; q < SSF_CHUNK; ++q) { re[q] = 0.0; im[q] = 0.0; }
            const double *px = sx + g * N, *pc = sc1 + g * N,
                         *ps = ss1 + g * N;
            for (int i = 0; i < N; ++i) {
                // three-term recurrence e^{i(m+1)t} = 2 cos t e^{imt}
                // - e^{i(m-1)t}, seeded exactly at m0: one DFMA per
                // component and mode; a rounding error is amplified by at
                // most the distance to the seed (<= 16)
                double cm, sm;
                sincospi((double) m0 * px[i], &sm, &cm);
                const double c1 = pc[i], s1 = ps[i], twoc = 2.0 * c1;
                double cp = fma(cm, c1, sm * s1);       // mode m0 - 1
                double sp = fma(sm, c1, -(cm * s1));
#pragma unroll
                for (int q = 0; q < SSF_CHUNK; ++q) {
                    re[q] += cm;
                    im[q] += sm;
                    double cn = fma(twoc, cm, -cp);
                    double sn = fma(twoc, sm, -sp);
                    cp = cm; sp = sm; cm = cn; sm = sn;
                }
            }
            double *o = a.out + (s * a.M + m0) * 3;
            const double *pv = nullptr;
            if (a.accumulate) {
                long long r = a.ref ? (long long) a.ref[s] : s;
                pv = a.prev + (r * a.M + m0) * 3;
            }
#pragma unroll
            for (int q = 0; q < SSF_CHUNK; ++q) {
                if (m0 + q < a.M) {
                    double v0 = fma(re[q], re[q], im[q] * im[q]);
                    double v1 = re[q], v2 = im[q];
                    if (pv) {
                        v0 += pv[3 * q]; v1 += pv[3 * q + 1];
                        v2 += pv[3 * q + 2];
                    }
                    o[3 * q] = v0; o[3 * q + 1] = v1; o[3 * q + 2] = v2;
                }
            }
        }
        __syncthreads();
    }
}

// Same result, cheaper seeds (the exact sincospi seed above is ~80
// instructions per particle and chunk, more than the 64 of the recurrence it
// starts).  Particles are taken in batches of P; for each one a stager thread
// evaluates e^{i theta} and e^{i 16 theta} exactly and walks the geometric
// sequence e^{i 16 j theta}, j = 0..nchunk-1, by complex multiplication (one
// rounding per step: <= nchunk ulp) into shared memory, from where the chunk
// threads pick their seed with one LDS.128.  Requires G * nchunk <= blockDim.
struct SsfStage {
    int G, nchunk, P;
};

__host__ __device__ inline size_t ssf_stage_smem(int G, int nchunk, int P)
{
    return (size_t) G * P * (nchunk + 1) * sizeof(double2);
}

__global__ void __launch_bounds__(SSF_THREADS)
ssf_eval_staged_kernel(SsfArgs a, SsfStage st)
{
    extern __shared__ __align__(16) double ssf_smem[];
    const long long W = a.W_dev ? (long long) *a.W_dev : a.W_host;
    const int N = a.N, G = st.G, nchunk = st.nchunk, P = st.P;
    double2 *seeds = reinterpret_cast<double2 *>(ssf_smem);  // [G][P][nchunk]
    double2 *cs = seeds + (size_t) G * P * nchunk;           // [G][P] (c1, s1)
    const int t = threadIdx.x;
    const bool worker = t < G * nchunk;
    const int g = worker ? t / nchunk : 0, j = worker ? t - g * nchunk : 0;
    const int m0 = j * SSF_CHUNK;
    for (long long s0 = (long long) blockIdx.x * G; s0 < W;
         s0 += (long long) gridDim.x * G) {
        const long long s = s0 + g;
        const bool live = worker && s < W;
        double re[SSF_CHUNK], im[SSF_CHUNK];
#pragma unroll
        for (int q = 0; q < SSF_CHUNK; ++q) { re[q] = 0.0; im[q] = 0.0; }
        for (int b0 = 0; b0 < N; b0 += P) {
            const int np = min(P, N - b0);
            __syncthreads();        // the previous batch has been consumed
            for (int e = t; e < G * np; e += blockDim.x) {
                const int gg = e / np, ii = e - gg * np;
                const long long ss = s0 + gg;
                if (ss >= W) continue;
                const long long r = a.ref ? (long long) a.ref[ss] : ss;
                const double x = a.confs[r * 2 * N + b0 + ii] * a.two_over_L;
                double s1, c1, sb, cb;
                sincospi(x, &s1, &c1);
                sincospi((double) SSF_CHUNK * x, &sb, &cb);
                cs[gg * P + ii] = make_double2(c1, s1);
                double2 *row = seeds + ((size_t) gg * P + ii) * nchunk;
                double cr = 1.0, ci = 0.0;
                for (int jj = 0; jj < nchunk; ++jj) {
                    row[jj] = make_double2(cr, ci);
                    const double nr = fma(cr, cb, -(ci * sb));
                    const double ni = fma(cr, sb, ci * cb);
                    cr = nr; ci = ni;
                }
            }
            __syncthreads();
            if (live) {
                const double2 *sd = seeds + (size_t) g * P * nchunk + j;
                const double2 *pc = cs + g * P;
                for (int ii = 0; ii < np; ++ii) {
                    const double2 e0 = sd[(size_t) ii * nchunk];
                    const double2 c = pc[ii];
                    double cm = e0.x, sm = e0.y;
                    const double twoc = 2.0 * c.x;
                    double cp = fma(cm, c.x, sm * c.y);      // mode m0 - 1
                    double sp = fma(sm, c.x, -(cm * c.y));
#pragma unroll
                    for (int q = 0; q < SSF_CHUNK; ++q) {
                        re[q] += cm;
                        im[q] += sm;
                        double cn = fma(twoc, cm, -cp);
                        double sn = fma(twoc, sm, -sp);
                        cp = cm; sp = sm; cm = cn; sm = sn;
                    }
                }
            }
        }
        if (live) {
            double *o = a.out + (s * a.M + m0) * 3;
            const double *pv = nullptr;
            if (a.accumulate) {
                long long r = a.ref ? (long long) a.ref[s] : s;
                pv = a.prev + (r * a.M + m0) * 3;
            }
#pragma unroll
            for (int q = 0; q < SSF_CHUNK; ++q) {
                if (m0 + q < a.M) {
                    double v0 = fma(re[q], re[q], im[q] * im[q]);
                    double v1 = re[q], v2 = im[q];
                    if (pv) {
                        v0 += pv[3 * q]; v1 += pv[3 * q + 1];
                        v2 += pv[3 * q + 2];
                    }
                    o[3 * q] = v0; o[3 * q + 1] = v1; o[3 * q + 2] = v2;
                }
            }
        }
    }
}

// Pure-estimator transport after the forward-walking window:
// out[s] = prev[ref[s]] (qmc_base/jastrow/dmc.py:441-447).
__global__ void rows_gather_kernel(const double *prev, const int *ref,
                                   const int *W_dev, int ncol, double *out)
{
    const long long W = *W_dev;
    const long long total = W * ncol;
    for (long long e = blockIdx.x * (long long) blockDim.x + threadIdx.x;
         e < total; e += (long long) gridDim.x * blockDim.x) {
        long long s = e / ncol;
        int c = (int) (e - s * ncol);
        out[e] = prev[(long long) ref[s] * ncol + c];
    }
}

// Column sums of rows [row_lo, row_hi) of a row-major [*][ncol] table:
// partial[blockIdx.x][col].  Rows are dealt to CTAs in contiguous slabs and
// summed in row order, so the result does not depend on scheduling.
struct RowRange {
    const int *lo_dev, *hi_dev;     // device scalars (null -> the host value)
    long long lo_host, hi_host;
};

__global__ void __launch_bounds__(256)
colsum_partial_kernel(const double *a, RowRange rr, int ncol, double *partial)
{
    const long long lo = rr.lo_dev ? (long long) *rr.lo_dev : rr.lo_host;
    const long long hi = rr.hi_dev ? (long long) *rr.hi_dev : rr.hi_host;
    const long long n = hi > lo ? hi - lo : 0;
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long r0 = lo + per * blockIdx.x;
    const long long r1 = (r0 + per < hi) ? r0 + per : hi;
    for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
        double acc = 0.0;
        long long r = r0;
        // eight loads in flight, added in row order (same sum as a plain loop)
        for (; r + 8 <= r1; r += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = a[(r + u) * ncol + c];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; r < r1; ++r) acc += a[r * ncol + c];
        partial[(long long) blockIdx.x * ncol + c] = acc;
    }
}

// out[c] = (base[c] * base_sign + sum_b partial[b][c]) * scale
__global__ void colsum_final_kernel(const double *partial, int nblk, int ncol,
                                    const double *base, double partial_sign,
                                    double scale, double *out)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    double acc = 0.0;
    int b = 0;
    for (; b + 8 <= nblk; b += 8) {     // eight loads in flight, same order
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            v[u] = partial[(long long) (b + u) * ncol + c];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; b < nblk; ++b) acc += partial[(long long) b * ncol + c];
    double v = partial_sign * acc;
    if (base) v += base[c];
    out[c] = v * scale;
}

// Bin index of a position: Python float floor division z // b (numba lowers
// `//` to: mod = fmod(z, b); div = (z - mod) / b; floor(div), rounded up when
// div - floor(div) > 0.5; mrbp_qmc/dmc.py:530,544).  For b > 0 that value is
// the exact floor of the real quotient z / b (fmod is exact and the 0.5 test
// repairs the rounding of the division), which needs no fmod:
// k = floor(z * (1/b)) is off by at most one, and the sign of the fma
// residual z - k b is exact and says which way.
__device__ __forceinline__ int floordiv_bin(double z, double b, double inv_b)
{
    double k = floor(z * inv_b);
    const double r = fma(-k, b, z);
    if (r < 0.0) k -= 1.0;
    else if (r >= b) k += 1.0;
    return (int) k;
}

// Per-slot histogram of the positions of the live walkers, plus the same
// counts into the global running histogram `total` (all slots), which is
// what lets the per-step sum over live slots be formed as
//   total - (rows of dead slots that were live earlier in the block).
// Counts are small integers held in doubles: the atomics are exact, hence
// order-independent.  The counts for `total` are first gathered per CTA in
// shared memory (32-bit atomics) so that the global array sees one atomic per
// touched bin and CTA.  hi_dev tracks the highest slot count seen.
__global__ void __launch_bounds__(256)
density_hist_kernel(const double *confs, const int *ref, const int *W_dev,
                    int N, int nbins, double bin_size, double *hist,
                    double *total, int *hi_dev, int use_smem)
{
    extern __shared__ unsigned int dens_smem[];
    const long long W = *W_dev;
    const double inv_bin = 1.0 / bin_size;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(hi_dev, (int) W);
    if (use_smem) {
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            dens_smem[b] = 0u;
        __syncthreads();
    }
    const long long n = W * N;
    for (long long e = blockIdx.x * (long long) blockDim.x + threadIdx.x;
         e < n; e += (long long) gridDim.x * blockDim.x) {
        long long s = e / N;
        int i = (int) (e - s * N);
        double z = confs[(long long) ref[s] * 2 * N + i];
        int b = floordiv_bin(z, bin_size, inv_bin);
        b = b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);     // quirk Q5: clamp
        atomicAdd(hist + s * nbins + b, 1.0);
        if (use_smem) atomicAdd(dens_smem + b, 1u);
        else atomicAdd(total + b, 1.0);
    }
    if (use_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
            unsigned int c = dens_smem[b];
            if (c) atomicAdd(total + b, (double) c);
        }
    }
}


// ---------------------------------------------------------------------------
// Density estimator without per-slot histograms.  In pure mode the reference
// ignores the genealogy (quirk Q2): slot s simply accumulates the counts of
// whatever walker occupies it, and step t sums the rows of the slots that are
// live at t; in mixed mode it does the same into two buffers by the parity of
// the step and never resets them within a block (quirk Q3).  Hence, with t'
// running over the recorded steps (of the same parity in mixed mode),
//   out(t) = [ sum over recorded steps t' <= t of ALL counts of t' ]
//            - [ counts of step t' in the slots W_t <= s < W_t' ]   (t' < t)
// The first term is one running histogram; the second touches only the few
// slots by which the population has shrunk since t'.  Instead of 8 B x bins
// per slot updated with N scattered read-modify-writes per walker and step,
// a step writes N 16-bit bin indices per walker, coalesced.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
density_list_kernel(const double *confs, const int *ref, const int *W_dev,
                    int N, int nbins, double bin_size, int cap,
                    unsigned short *lists,      // [cap][N] of this step
                    int *W_rec,                 // population of this step
                    double *total, int use_smem)
{
    extern __shared__ unsigned int dens_smem[];
    const long long W = *W_dev;
    const double inv_bin = 1.0 / bin_size;
    if (blockIdx.x == 0 && threadIdx.x == 0) *W_rec = (int) W;
    if (use_smem) {
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            dens_smem[b] = 0u;
        __syncthreads();
    }
    const long long n = W * N;
    for (long long e = blockIdx.x * (long long) blockDim.x + threadIdx.x;
         e < n; e += (long long) gridDim.x * blockDim.x) {
        long long s = e / N;
        int i = (int) (e - s * N);
        double z = confs[(long long) ref[s] * 2 * N + i];
        int b = floordiv_bin(z, bin_size, inv_bin);
        b = b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);     // quirk Q5: clamp
        lists[e] = (unsigned short) b;
        if (use_smem) atomicAdd(dens_smem + b, 1u);
        else atomicAdd(total + b, 1.0);
    }
    if (use_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
            unsigned int c = dens_smem[b];
            if (c) atomicAdd(total + b, (double) c);
        }
    }
}

// CTA t': counts of recorded step t' in the slots that have died since
// (exact integer-valued atomics: order-independent).
__global__ void __launch_bounds__(256)
density_corr_kernel(const unsigned short *lists, const int *W_rec,
                    const int *W_dev, int N, long long step_stride,
                    int first, int every, double *corr)
{
    // recorded steps first, first + every, ...: all of them in pure mode,
    // those of the current parity in mixed mode
    const long long tp = first + (long long) every * blockIdx.x;
    const long long Wt = *W_dev, Wp = W_rec[tp];
    if (Wp <= Wt) return;
    const unsigned short *l = lists + tp * step_stride;
    for (long long e = Wt * N + blockIdx.y * blockDim.x + threadIdx.x;
         e < Wp * N; e += (long long) gridDim.y * blockDim.x)
        atomicAdd(corr + l[e], 1.0);
}

__global__ void density_out_kernel(const double *total, double *corr,
                                   int nbins, double scale, double *out)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    out[b] = (total[b] - corr[b]) * scale;
    corr[b] = 0.0;
}

// ---------------------------------------------------------------------------
// One-body density matrix  g1(sz) = (1/N) sum_i Psi(.., z_i + sz, ..) / Psi
// of fixed configurations (qmc_base/jastrow/model.py:859-965; the gufunc of
// PhysicalFuncs :1069-1091 broadcasts it over offsets and configurations).
//
// One thread owns one (configuration, offset) item and walks the shifted
// particle i and its partners j itself, so no cross-thread reduction is
// needed; the items of a CTA span a few configurations whose column tables
// (the same far table / four pre-rotated near variants as the step kernel)
// sit in shared memory, read as warp-wide broadcasts.  ln f2 is never taken
// per pair: |sin| (far) and |cos| (near) factors are multiplied up with
// exponent renormalisation and one log per row closes the product.  The
// unshifted row products are formed once per particle and shared by every
// offset of the configuration.
// ---------------------------------------------------------------------------
constexpr int OBD_ROW_DOUBLES = 15;  // per particle: A 2, V 8, z, ln f1, lf, ln, nnear

// doubles per configuration slot, even so that the double2 tables stay
// 16-byte aligned
__host__ __device__ inline size_t obd_slot_doubles(int N)
{
    return ((size_t) OBD_ROW_DOUBLES * N + 1) & ~(size_t) 1;
}

struct ObdArgs {
    const double *confs;    // [nconf][2][N]
    long long nconf;
    const double *offsets;  // [S]
    int S;
    double *out;            // [nconf][S]
    int per_conf;           // 1: a CTA's items belong to ONE configuration
                            // (large N: tables of one configuration at a time)
};

// Products over the partners j != iskip of a particle with tables
// (sa, ca, su, cu): lf = ln prod |sin(a - a_j)| / gamma_f over far pairs,
// ln = ln prod |cos(u - u_j')| over near pairs, nnear = number of near pairs.
__device__ __forceinline__ void obd_row(const DevModel &M, const double2 *A,
                                        const double2 *V, int iskip,
                                        double sa, double ca, double su,
                                        double cu, double &lf, double &ln,
                                        int &nnear)
{
    const int N = M.nop;
    const double s_m = M.s_m_scaled;
    double pf = 1.0, pn = 1.0;
    int ef = 0, en = 0, nn = 0;
    for (int j0 = 0; j0 < N; j0 += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + q;
            const bool ok = (j < N) && (j != iskip);
            const int jj = j < N ? j : N - 1;
            const double2 a = A[jj];
            const double den_f = fma(sa, a.y, -(ca * a.x));
            const double cosd = fma(ca, a.y, sa * a.x);
            // bit 1 = unwrapped (cos >= 0), bit 0 = sigma > 0
            const unsigned hc = ~(unsigned) __double2hiint(cosd);
            const unsigned hd = (unsigned) __double2hiint(den_f);
            const unsigned v = ((hc >> 31) << 1) + ((hc ^ hd) >> 31);
            const double2 vv = V[jj * 4 + v];
            const double den_n = fma(cu, vv.y, su * vv.x);
            const bool near = fabs(den_f) < s_m;
            pf *= (ok && !near) ? fabs(den_f) : 1.0;
            pn *= (ok && near) ? fabs(den_n) : 1.0;
            nn += (ok && near) ? 1 : 0;
        }
        renorm(pf, ef);
        renorm(pn, en);
    }
    lf = log(pf) + ef * LN2;
    ln = log(pn) + en * LN2;
    nnear = nn;
}

__global__ void obd_kernel(DevModel M, ObdArgs a)
{
    extern __shared__ __align__(16) double obd_smem[];
    const int N = M.nop, S = a.S;
    const long long wtot = a.nconf * S;
    long long w0, wl;       // first and last item of this CTA
    if (a.per_conf) {
        const int nchunk = (S + blockDim.x - 1) / blockDim.x;
        const long long c = blockIdx.x / nchunk;
        const int s0 = (int) (blockIdx.x - c * nchunk) * blockDim.x;
        if (c >= a.nconf) return;
        w0 = c * S + s0;
        wl = c * S + (s0 + (int) blockDim.x < S ? s0 + (int) blockDim.x : S) - 1;
    } else {
        w0 = (long long) blockIdx.x * blockDim.x;
        if (w0 >= wtot) return;
        wl = (w0 + blockDim.x < wtot ? w0 + blockDim.x : wtot) - 1;
    }
    const long long c0 = w0 / S;
    const int nc = (int) (wl / S - c0) + 1;
    const size_t slot = obd_slot_doubles(N);
    // tables of the configurations this CTA touches
    for (int e = threadIdx.x; e < nc * N; e += blockDim.x) {
        const int g = e / N, i = e - g * N;
        double *base = obd_smem + g * slot;
        double2 *A = reinterpret_cast<double2 *>(base);
        double2 *V = A + N;
        double *zs = base + 10 * (size_t) N;
        const double z = a.confs[(c0 + g) * 2 * N + i];
        zs[i] = z;
        if (!M.is_ideal) {
            double sa, ca, su, cu;
            particle_tables(M, recast(z, 0.0, M.L), sa, ca, su, cu);
            A[i] = make_double2(sa * M.inv_gam, ca * M.inv_gam);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int w = (v & 2) ? 0 : 1;
                const double cp = M.cpsi[w];
                const double sp = (v & 1) ? M.spsi[w] : -M.spsi[w];
                V[i * 4 + v] = make_double2(fma(su, cp, cu * sp),
                                            fma(cu, cp, -(su * sp)));
            }
        }
    }
    __syncthreads();
    // unshifted row products and ln f1, once per particle
    for (int e = threadIdx.x; e < nc * N; e += blockDim.x) {
        const int g = e / N, i = e - g * N;
        double *base = obd_smem + g * slot;
        const double2 *A = reinterpret_cast<const double2 *>(base);
        const double2 *V = A + N;
        double *zs = base + 10 * (size_t) N;
        const double z = zs[i];
        double ln1 = 0.0, lf = 0.0, ln = 0.0;
        int nnear = 0;
        if (!M.is_free) ln1 = one_body<true>(M, z).lnf;
        if (!M.is_ideal) {
            double sa, ca, su, cu;
            particle_tables(M, recast(z, 0.0, M.L), sa, ca, su, cu);
            obd_row(M, A, V, i, sa, ca, su, cu, lf, ln, nnear);
        }
        zs[N + i] = ln1;
        zs[2 * N + i] = lf;
        zs[3 * N + i] = ln;
        reinterpret_cast<int *>(zs + 4 * N)[i] = nnear;
    }
    __syncthreads();
    const long long w = w0 + threadIdx.x;
    if (w > wl) return;
    const long long c = w / S;
    const int s = (int) (w - c * S);
    const double *base = obd_smem + (size_t) (c - c0) * slot;
    const double2 *A = reinterpret_cast<const double2 *>(base);
    const double2 *V = A + N;
    const double *zs = base + 10 * (size_t) N;
    const int *cnt = reinterpret_cast<const int *>(zs + 4 * N);
    const double sz = a.offsets[s];
    double acc = 0.0;
    // the reference returns 0 (not exp(0)) per particle for the free ideal
    // gas (qmc_base/jastrow/model.py:891-892)
    if (!(M.is_free && M.is_ideal)) {
        for (int i = 0; i < N; ++i) {
            const double zsft = zs[i] + sz;
            double lnr = 0.0;
            if (!M.is_free)
                lnr = one_body<true>(M, zsft).lnf - zs[N + i];
            if (!M.is_ideal) {
                double sa, ca, su, cu, lf, ln;
                int nnear;
                particle_tables(M, recast(zsft, 0.0, M.L), sa, ca, su, cu);
                obd_row(M, A, V, i, sa, ca, su, cu, lf, ln, nnear);
                const int dn = nnear - cnt[i];
                lnr += M.beta * ((lf - zs[2 * N + i]) - dn * M.ln_gam)
                       + (ln - zs[3 * N + i]) + dn * M.ln_am;
            }
            acc += exp(lnr);
        }
    }
    a.out[w] = acc / N;
}


// ---------------------------------------------------------------------------
// rho_k of fixed configurations at ARBITRARY momenta (the gufunc
// PhysicalFuncs.fourier_density, qmc_base/jastrow/model.py:1093-1122, takes
// any kz_set; the samplers only use k_m = 2 pi m / L, which ssf_eval_kernel
// covers).  One thread per (configuration, momentum); out [nconf][nk][2] =
// (Re, Im) = (sum cos kz z_i, sum sin kz z_i) summed in particle order.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fdk_general_kernel(const double *confs, long long nconf, int N,
                   const double *kz, int nk, double *out)
{
    const long long total = nconf * nk;
    for (long long e = blockIdx.x * (long long) blockDim.x + threadIdx.x;
         e < total; e += (long long) gridDim.x * blockDim.x) {
        const long long b = e / nk;
        const double k = kz[e - b * nk];
        const double *z = confs + b * 2 * N;
        double sc = 0.0, ss = 0.0;
        for (int i = 0; i < N; ++i) {
            double s, c;
            sincos(k * z[i], &s, &c);
            sc += c;
            ss += s;
        }
        out[2 * e] = sc;
        out[2 * e + 1] = ss;
    }
}

// ---------------------------------------------------------------------------
// Correlated-sampling objective of the wave-function optimiser
// (CSWFOptimizer.weighed_variance, qmc_base/jastrow/model.py:1147-1165):
//   w_c = exp(2 (ln|Psi_trial| - ln|Psi_0|)_c - max),  E_ref = <E>_w,
//   variance = <(E - E_ref)^2>_w.
// One CTA, fixed-order tree reductions: the result does not depend on
// scheduling.  out = {variance, E_ref, sum w, sum w^2}.
// ---------------------------------------------------------------------------
constexpr int CS_THREADS = 1024;

template <typename Op>
__device__ __forceinline__ double cs_block_reduce(double v, double *red, Op op)
{
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = CS_THREADS / 2; s > 0; s >>= 1) {
        if ((int) threadIdx.x < s)
            red[threadIdx.x] = op(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    double r = red[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(CS_THREADS)
cs_variance_kernel(const double *lnpsi, const double *lnpsi0,
                   const double *energy, long long n, double *out)
{
    __shared__ double red[CS_THREADS];
    auto fmaxop = [](double a, double b) { return a > b ? a : b; };
    auto addop = [](double a, double b) { return a + b; };
    double m = -INFINITY;
    for (long long c = threadIdx.x; c < n; c += CS_THREADS)
        m = fmaxop(m, 2.0 * (lnpsi[c] - lnpsi0[c]));
    m = cs_block_reduce(m, red, fmaxop);
    double sw = 0.0, swe = 0.0, sw2 = 0.0;
    for (long long c = threadIdx.x; c < n; c += CS_THREADS) {
        double w = exp(2.0 * (lnpsi[c] - lnpsi0[c]) - m);
        sw += w;
        swe = fma(w, energy[c], swe);
        sw2 = fma(w, w, sw2);
    }
    sw = cs_block_reduce(sw, red, addop);
    swe = cs_block_reduce(swe, red, addop);
    sw2 = cs_block_reduce(sw2, red, addop);
    const double eref = swe / sw;
    double sv = 0.0;
    for (long long c = threadIdx.x; c < n; c += CS_THREADS) {
        double w = exp(2.0 * (lnpsi[c] - lnpsi0[c]) - m);
        double d = energy[c] - eref;
        sv = fma(w, d * d, sv);
    }
    sv = cs_block_reduce(sv, red, addop);
    if (threadIdx.x == 0) {
        out[0] = sv / sw;
        out[1] = eref;
        out[2] = sw;
        out[3] = sw2;
    }
}

}  // namespace qmcb
