// Batched Metropolis VMC: one chain per walker group of a CTA, the whole
// block of `ns` steps inside one launch (chains never interact).
//
// Reference (paths relative to src/phd_qmclib/ of PhD-QMCLib), per chain:
//   states_generator / blocks          qmc_base/vmc.py:557-648, 670-770
//   all-particle uniform proposal      qmc_base/jastrow/vmc.py:201-226,
//                                      mrbp_qmc/vmc.py:206-235
//   energy / S(k) on acceptance only   qmc_base/jastrow/vmc.py:229-351
//   momenta k_m = m 2 pi / L, m >= 0   mrbp_qmc/vmc.py:242-271
#pragma once
#include "qmcb_kernels.cuh"

namespace qmcb {

constexpr int VMC_MB = 10;      // S(k) modes per batch (2 * VMC_MB * nb <= 20 * nbp)
constexpr int VMC_RESEED = 60;  // exact phase re-seed period (modes; multiple of VMC_MB)

static_assert(VMC_RESEED % VMC_MB == 0, "re-seed at batch boundaries");

struct VmcState {
    double *confs;      // [C][2][N]   row 0 positions, row 1 drift
    double *lnpsi;      // [C]
    double *eprev;      // [C]         energy of the last accepted state
    double *ssfprev;    // [C][M][3]   S(k) parts of the last accepted state
};

struct VmcArgs {
    long long nchains, ns;
    long long gstep0;       // generator step index of the first RNG draw
    long long chain_offset;
    int first;              // first yielded state = the initial one, ACCEPTED
    int M;
    int proposal;           // 0 uniform of width `spread`, 1 gaussian sigma
    uint64_t seed;
    double spread, z_min, size, two_over_L;
    double *out_lnpsi, *out_energy;     // [C][ns] or null
    unsigned char *out_stat;            // [C][ns] or null
    double *out_ssf;                    // [C][ns][M][3] or null
    double *accept_rate;                // [C] or null
    double *sum_energy;                 // [C][2] (sum e, sum e^2) or null
    double *sum_ssf;                    // [C][M][3] or null
    double *out_confs;                  // [C][ns][2][N] or null: every state
                                        // of the chain (as_chain,
                                        // qmc_base/vmc.py:773-902)
};

// FAST: node-table transcendentals (TrigTab); needs every position in
// [0, L]: the recast interval equal to the supercell and initial positions
// inside it (checked by the host).
// 1 in *flag when some position of row 0 of `confs` [C][2][N] is outside
// [0, L] (NaN included): the table path of the block kernel then stays off.
__global__ void positions_outside_kernel(const double *__restrict__ confs,
                                         long long nchains, int nop, double L,
                                         int *flag)
{
    const long long total = nchains * nop;
    bool bad = false;
    for (long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
         t < total; t += (long long) gridDim.x * blockDim.x) {
        const long long c = t / nop;
        const double z = confs[c * 2 * nop + (t - c * nop)];
        bad |= !(z >= 0.0 && z <= L);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0)
        atomicOr(flag, 1);
}

template <bool FAST>
__global__ void __launch_bounds__(256, 2)
vmc_block_kernel(const __grid_constant__ DevModel M, GroupGeom geom,
                 VmcState S, VmcArgs a)
{
    GroupSmem sm = group_smem(geom);
    GroupIdx x = group_index(M, geom);
    const int N = M.nop, nb = M.nb;
    const bool vec_ok = (N % 2) == 0;
    const int nvalid = min(TB, N - TB * x.I);
    // accepted-chain masks of two consecutive steps (compacted S(k) phases)
    __shared__ unsigned take_mask[2];
    const bool compact = geom.G <= 32 && a.M > 0;
    if (threadIdx.x < 2) take_mask[threadIdx.x] = 0u;
    __syncthreads();
    for (long long base = (long long) blockIdx.x * geom.G; base < a.nchains;
         base += (long long) gridDim.x * geom.G) {
        const long long c = base + x.g;
        const bool active = x.in_group && c < a.nchains;
        double z[TB] = {}, Fcur[TB] = {};
        double ln_cur = 0.0, e_prev = 0.0, se = 0.0, se2 = 0.0;
        long long nacc = 0;
        int nstay = 0;      // steps the current state has not yet been added
                            // to the S(k) block sums for
        if (active) {
            load4(S.confs + c * 2 * N, x.I, nvalid, vec_ok, z);
            load4(S.confs + c * 2 * N + N, x.I, nvalid, vec_ok, Fcur);
            ln_cur = S.lnpsi[c];
            e_prev = S.eprev[c];
        }
        const uint32_t gc = (uint32_t) (a.chain_offset + c);
        // S(k) partials, two buffers of [VMC_MB][nb], reuse the walker's pair
        // tables, which are dead between two evaluations (2 * VMC_MB * nb <=
        // 20 * nbp double2)
        double2 *part = reinterpret_cast<double2 *>(sm.tab(x.g));
        for (long long st = 0; st < a.ns; ++st) {
            const bool ini = a.first && st == 0;
            const long long gs = a.gstep0 + st - (a.first ? 1 : 0);
            double zp[TB];
#pragma unroll
            for (int q = 0; q < TB; ++q) zp[q] = z[q];
            if (active && !ini) {
                double u[TB];
                if (a.proposal == 1) {
                    {
                        double n4[4];
                        rng_normal4<FAST>(M.tt, a.seed, gc,
                                          (uint32_t) ((TB * x.I) / 4),
                                          (uint32_t) gs, STREAM_VMC_MOVE, n4);
#pragma unroll
                        for (int q = 0; q < TB; ++q)
                            u[q] = n4[((TB * x.I) % 4 + q) & 3];
                    }
#pragma unroll
                    for (int q = 0; q < TB; ++q)
                        zp[q] = recast(z[q] + a.spread * u[q], a.z_min,
                                       a.size);
                } else {
                    {
                        double u4[4];
                        rng_uniform4(a.seed, gc, (uint32_t) ((TB * x.I) / 4),
                                     (uint32_t) gs, STREAM_VMC_MOVE, u4);
#pragma unroll
                        for (int q = 0; q < TB; ++q)
                            u[q] = u4[((TB * x.I) % 4 + q) & 3];
                    }
#pragma unroll
                    for (int q = 0; q < TB; ++q)
                        zp[q] = recast(z[q] + (u[q] - 0.5) * a.spread,
                                       a.z_min, a.size);
                }
            }
            EvalOut o;
            group_eval<true, true, true, FAST>(M, sm, x.g, x.I, active, zp,
                                               nvalid, o);
            bool take = false;
            if (active) {
                if (ini) {
                    take = true;
                } else {
                    double ua, ub;
                    rng_uniform2(a.seed, gc, 0u, (uint32_t) gs,
                                 STREAM_VMC_ACCEPT, ua, ub);
                    // qmc_base/vmc.py:636
                    take = o.lnpsi > 0.5 * log(ua) + ln_cur;
                }
                if (take) {
#pragma unroll
                    for (int q = 0; q < TB; ++q) {
                        z[q] = zp[q];
                        Fcur[q] = o.F[q];
                    }
                    if (!ini) ln_cur = o.lnpsi;
                    e_prev = o.energy;
                    ++nacc;
                }
                se += e_prev;
                se2 = fma(e_prev, e_prev, se2);
                if (a.out_confs) {
                    double *oc = a.out_confs + ((c * a.ns + st) * 2) * N;
                    store4(oc, x.I, nvalid, vec_ok, z);
                    store4(oc + N, x.I, nvalid, vec_ok, Fcur);
                }
                if (x.I == 0) {
                    if (a.out_lnpsi) a.out_lnpsi[c * a.ns + st] = ln_cur;
                    if (a.out_energy) a.out_energy[c * a.ns + st] = e_prev;
                    if (a.out_stat) a.out_stat[c * a.ns + st] = take ? 1 : 0;
                }
            }
            if (a.M > 0) {
                // rho_k of the accepted configuration: every thread advances
                // e^{i k_m z} of its own particles mode by mode; the sum over
                // the chain's threads goes through shared memory, VMC_MB
                // modes per barrier pair.
                // Three-term recurrence e^{i(m+1)t} = 2 cos t e^{imt} -
                // e^{i(m-1)t} (one DFMA per component and mode), exact
                // re-seed every VMC_RESEED modes; the phases of padding
                // particles are zero and stay zero, so the sums need no
                // predicate.
                // The phase work is COMPACTED: with ~half of the proposals
                // rejected, a warp of chain threads would run it half empty.
                // The accepted chains park their new positions in the state
                // array and the first (accepted chains) x nb threads of the
                // CTA take one (chain, particle block) unit each: whole warps
                // work, the others skip.  Pairing by rank from a mask word.
                const int par = (int) (st & 1);
                bool prod = take;
                int ug = x.g, ub = x.I, unv = nvalid;
                double zz[TB];
#pragma unroll
                for (int q = 0; q < TB; ++q) zz[q] = z[q];
                if (compact) {
                    if (active && take) {
                        store4(S.confs + c * 2 * N, x.I, nvalid, vec_ok, z);
                        if (x.I == 0) atomicOr(&take_mask[par], 1u << x.g);
                    }
                    __syncthreads();
                    const unsigned tk = take_mask[par];
                    const int k = (int) threadIdx.x / nb;
                    ub = (int) threadIdx.x - k * nb;
                    prod = k < __popc(tk);
                    ug = prod ? (int) __fns(tk, 0u, k + 1) : 0;
                    unv = min(TB, N - TB * ub);
                    if (prod)
                        load4(S.confs + (base + ug) * 2 * N, ub, unv, vec_ok,
                              zz);
                    if (threadIdx.x == 0) take_mask[par ^ 1] = 0u;
                }
                double2 *upart = reinterpret_cast<double2 *>(sm.tab(ug));
                double pc[TB], ps[TB], pcp[TB], psp[TB], twoc[TB], xs[TB];
#pragma unroll
                for (int q = 0; q < TB; ++q) {
                    xs[q] = zz[q] * a.two_over_L;
                    pc[q] = 0.0; ps[q] = 0.0; pcp[q] = 0.0; psp[q] = 0.0;
                    twoc[q] = 2.0;
                    if (prod && q < unv) {
                        double s1, c1;
                        sincospi(xs[q], &s1, &c1);
                        pc[q] = 1.0;                    // mode 0
                        pcp[q] = c1; psp[q] = -s1;      // mode -1
                        twoc[q] = 2.0 * c1;
                    }
                }
                // The partials are double-buffered: one CTA barrier per
                // batch.  The block sums are accumulated with fire-and-forget
                // reductions (the adds to one address are a barrier apart:
                // their order is fixed).
                int buf = 0;
                for (int m0 = 0; m0 < a.M; m0 += VMC_MB, buf ^= 1) {
                    double2 *pb = part + buf * (nb * VMC_MB);
                    if (prod) {
                        double2 *upb = upart + buf * (nb * VMC_MB);
                        if (m0 > 0 && (m0 % VMC_RESEED) == 0) {
#pragma unroll
                            for (int q = 0; q < TB; ++q)
                                if (q < unv) {
                                    double s1, c1;
                                    sincospi(xs[q], &s1, &c1);
                                    sincospi((double) m0 * xs[q], &ps[q],
                                             &pc[q]);
                                    pcp[q] = fma(pc[q], c1, ps[q] * s1);
                                    psp[q] = fma(ps[q], c1, -(pc[q] * s1));
                                }
                        }
#pragma unroll
                        for (int j = 0; j < VMC_MB; ++j) {
                            double re = 0.0, im = 0.0;
#pragma unroll
                            for (int q = 0; q < TB; ++q) {
                                re += pc[q]; im += ps[q];
                                const double cn = fma(twoc[q], pc[q], -pcp[q]);
                                const double sn = fma(twoc[q], ps[q], -psp[q]);
                                pcp[q] = pc[q]; psp[q] = ps[q];
                                pc[q] = cn; ps[q] = sn;
                            }
                            upb[j * nb + ub] = make_double2(re, im);
                        }
                    }
                    // Block sums, lazily: a state that stays for n steps
                    // adds n times its parts when it is replaced (or at the
                    // end of the block), so a rejected proposal costs no
                    // memory traffic here.  The replaced state's parts are
                    // fetched ahead of the barrier, in flight while the
                    // phases of this batch are produced.
                    double o0 = 0.0, o1 = 0.0, o2 = 0.0;
                    const bool flush = take && nstay > 0 && a.sum_ssf;
                    if (flush && x.I < VMC_MB && m0 + x.I < a.M) {
                        const double *sp = S.ssfprev
                            + (c * a.M + m0 + x.I) * 3;
                        o0 = sp[0]; o1 = sp[1]; o2 = sp[2];
                    }
                    if (a.out_ssf && active && !take) {
                        for (int j = x.I; j < VMC_MB && m0 + j < a.M;
                             j += nb) {
                            const int m = m0 + j;
                            const double *sp = S.ssfprev + (c * a.M + m) * 3;
                            double *os = a.out_ssf
                                + ((c * a.ns + st) * a.M + m) * 3;
                            os[0] = sp[0]; os[1] = sp[1]; os[2] = sp[2];
                        }
                    }
                    __syncthreads();
                    for (int j = x.I; take && j < VMC_MB && m0 + j < a.M;
                         j += nb) {
                        const int m = m0 + j;
                        double *sp = S.ssfprev + (c * a.M + m) * 3;
                        if (flush) {
                            if (j != x.I) { o0 = sp[0]; o1 = sp[1]; o2 = sp[2]; }
                            double *ss = a.sum_ssf + (c * a.M + m) * 3;
                            const double w = (double) nstay;
                            atomicAdd(ss, w * o0); atomicAdd(ss + 1, w * o1);
                            atomicAdd(ss + 2, w * o2);
                        }
                        // two interleaved running sums (fixed order)
                        const double2 *pj = pb + j * nb;
                        double re = 0.0, im = 0.0, re1 = 0.0, im1 = 0.0;
                        int t = 0;
                        for (; t + 1 < nb; t += 2) {
                            const double2 p0 = pj[t], p1 = pj[t + 1];
                            re += p0.x; im += p0.y;
                            re1 += p1.x; im1 += p1.y;
                        }
                        if (t < nb) { re += pj[t].x; im += pj[t].y; }
                        re += re1; im += im1;
                        const double v0 = fma(re, re, im * im);
                        sp[0] = v0; sp[1] = re; sp[2] = im;
                        if (a.out_ssf) {
                            double *os = a.out_ssf
                                + ((c * a.ns + st) * a.M + m) * 3;
                            os[0] = v0; os[1] = re; os[2] = im;
                        }
                    }
                }
                if (active) nstay = take ? 1 : nstay + 1;
                // the partials alias the pair tables of the next evaluation
                __syncthreads();
            }
        }
        if (active) {
            // the last state's share of the block sums
            if (a.M > 0 && a.sum_ssf && nstay > 0) {
                const double w = (double) nstay;
                for (int m = x.I; m < a.M; m += nb) {
                    const double *sp = S.ssfprev + (c * a.M + m) * 3;
                    double *ss = a.sum_ssf + (c * a.M + m) * 3;
                    atomicAdd(ss, w * sp[0]); atomicAdd(ss + 1, w * sp[1]);
                    atomicAdd(ss + 2, w * sp[2]);
                }
            }
            store4(S.confs + c * 2 * N, x.I, nvalid, vec_ok, z);
            store4(S.confs + c * 2 * N + N, x.I, nvalid, vec_ok, Fcur);
            if (x.I == 0) {
                S.lnpsi[c] = ln_cur;
                S.eprev[c] = e_prev;
                if (a.accept_rate)
                    a.accept_rate[c] = (double) nacc / (double) a.ns;
                if (a.sum_energy) {
                    a.sum_energy[2 * c] = se;
                    a.sum_energy[2 * c + 1] = se2;
                }
            }
        }
        if (threadIdx.x < 2) take_mask[threadIdx.x] = 0u;   // next pass
        __syncthreads();
    }
}

}  // namespace qmcb
