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
        for (long long r = r0; r < r1; ++r) acc += a[r * ncol + c];
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
    for (int b = 0; b < nblk; ++b) acc += partial[(long long) b * ncol + c];
    double v = partial_sign * acc;
    if (base) v += base[c];
    out[c] = v * scale;
}

// Python float floor division z // b (numba lowers `//` to this), as an
// integer bin index (mrbp_qmc/dmc.py:530,544).
__device__ __forceinline__ int py_floordiv_bin(double z, double b)
{
    double mod = fmod(z, b);
    double div = (z - mod) / b;
    if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return (int) fl;
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
        int b = py_floordiv_bin(z, bin_size);
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

}  // namespace qmcb
