// Kernels of the B200 walker-ensemble engine (model evaluation, DMC step,
// branching, population control).  See qmcb_dev.cuh for the math.
#pragma once
#include "qmcb_dev.cuh"

namespace qmcb {

// Thread (g, I) of a walker-group CTA.
struct GroupIdx {
    int g, I;
    bool in_group;      // thread maps to a (walker, block) pair at all
};

__device__ __forceinline__ GroupIdx group_index(const DevModel &M,
                                                const GroupGeom &geom)
{
    GroupIdx x;
    const int t = threadIdx.x, G = geom.G;
    if (geom.interleave) {
        x.I = t / G;
        x.g = t - x.I * G;
        x.in_group = x.I < M.nb;
    } else {
        x.g = t / M.nb;
        x.I = t - x.g * M.nb;
        x.in_group = x.g < G;
    }
    if (!x.in_group) { x.g = 0; x.I = 0; }
    return x;
}

__device__ __forceinline__ GroupSmem group_smem(const GroupGeom &geom)
{
    extern __shared__ __align__(16) double qmcb_smem[];
    GroupSmem sm;
    sm.base = qmcb_smem;
    sm.nbp = geom.nbp;
    sm.kc = geom.kc;
    sm.G = geom.G;
    sm.tab_stride = geom.tab_stride;
    sm.q_stride = geom.q_stride;
    return sm;
}

// Load / store the thread's 4 values of one row of a (2, N) configuration.
__device__ __forceinline__ void load4(const double *row, int I, int nvalid,
                                      bool vec_ok, double (&x)[TB])
{
    if (vec_ok && nvalid == TB) {
        const double2 *p = reinterpret_cast<const double2 *>(row + TB * I);
#pragma unroll
        for (int c = 0; c < TB; c += 2) {
            const double2 a = p[c / 2];
            x[c] = a.x; x[c + 1] = a.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < TB; ++c)
            x[c] = (c < nvalid) ? row[TB * I + c] : 0.0;
    }
}

__device__ __forceinline__ void store4(double *row, int I, int nvalid,
                                       bool vec_ok, const double (&x)[TB])
{
    if (vec_ok && nvalid == TB) {
        double2 *p = reinterpret_cast<double2 *>(row + TB * I);
#pragma unroll
        for (int c = 0; c < TB; c += 2) p[c / 2] = make_double2(x[c], x[c + 1]);
    } else {
#pragma unroll
        for (int c = 0; c < TB; ++c)
            if (c < nvalid) row[TB * I + c] = x[c];
    }
}

// ---------------------------------------------------------------------------
// K1 / K9: lnPsi, E_L, drift of a batch of configurations.
// Output layout selectable: separate arrays (model_eval) or a DMC state
// buffer (init: confs[s] = (z, F), energy[s], weight[s] = 1).
// ---------------------------------------------------------------------------
struct EvalArgs {
    const double *confs;    // [nconf][2][N]
    long long nconf;
    double *lnpsi;          // [nconf] or null
    double *energy;         // [nconf] or null
    double *drift;          // [nconf][N] or null
    double *state_confs;    // [cap][2][N] or null: write (z, F)
    double *state_weight;   // [cap] or null: write 1
    double *slot_energy;    // [cap] or null: write E
};

template <bool LN, bool EF>
__global__ void __launch_bounds__(256)
model_eval_kernel(const __grid_constant__ DevModel M, GroupGeom geom,
                  EvalArgs a)
{
    GroupSmem sm = group_smem(geom);
    GroupIdx x = group_index(M, geom);
    const int N = M.nop;
    const bool vec_ok = (N % 2) == 0;
    for (long long base = (long long) blockIdx.x * geom.G; base < a.nconf;
         base += (long long) gridDim.x * geom.G) {
        long long b = base + x.g;
        bool active = x.in_group && b < a.nconf;
        int nvalid = min(TB, N - TB * x.I);
        double z[TB] = {};
        if (active) {
            load4(a.confs + b * 2 * N, x.I, nvalid, vec_ok, z);
            if (a.state_confs)
                store4(a.state_confs + b * 2 * N, x.I, nvalid, vec_ok, z);
        }
        EvalOut o;
        group_eval<LN, EF, false>(M, sm, x.g, x.I, active, z, nvalid, o);
        if (active) {
            if (EF && a.drift)
                store4(a.drift + b * N, x.I, nvalid, vec_ok, o.F);
            if (a.state_confs)
                store4(a.state_confs + b * 2 * N + N, x.I, nvalid, vec_ok,
                       o.F);
            if (x.I == 0) {
                if (LN && a.lnpsi) a.lnpsi[b] = o.lnpsi;
                if (EF && a.energy) a.energy[b] = o.energy;
                if (EF && a.slot_energy) a.slot_energy[b] = o.energy;
                if (a.state_weight) a.state_weight[b] = 1.0;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// DMC control block (device memory).  Everything a time step needs lives
// here, so a step is a fixed sequence of launches with constant arguments.
// ---------------------------------------------------------------------------
struct DmcCtl {
    long long step;             // time steps completed since init
    long long capacity_hits;
    long long total_children;   // before truncation at capacity
    int W_prev;                 // live walkers of the population to branch
    int W;                      // live walkers after branching (this step)
    double eref[2];             // eref[step & 1] drives step `step`
    double tot_e, tot_w;        // running totals (qmc_base/dmc.py:733-765)
    double red[2];              // {sum E_parents, W}: local, then global
    double last_energy, last_weight, last_accum;
    double W_global;            // global live walkers of the last step
    unsigned int done_fill;     // CTAs of branch_fill_kernel that finished
    long long block_step0;      // `step` at the start of the block in flight
    long long tcur;             // time step the step kernel in flight works
                                // on (set by branch_fill_kernel, so that the
                                // population control may run next to it)
    long long pos_base[2];      // global position (index in the ordered
                                // concatenation of all ranks' slabs) of local
                                // slot 0 of the population step t branches
                                // from: pos_base[t & 1].  It keys the RNG, so
                                // a sharded run draws the numbers of the
                                // single-rank run
};

struct DmcBufs {
    double *confs[2];           // [cap][2][N]
    double *energy[2];          // [cap]
    double *weight[2];          // [cap]
    double *slot_energy;        // [cap]  persistent per-slot array (quirk Q1)
    int *ref;                   // [cap]  child slot -> parent slot
    int *cnt;                   // [cap]  clone counts
    long long *blocksum;        // [nblk]
    double *epart;              // [nblk]
    DmcCtl *ctl;
    int cap;
    int nblk;
};

struct DmcConsts {
    double dt, sigma, z_min, size, nwc_over_dt, target;
    uint64_t seed;
    long long slot_offset;      // global index of local slot 0 at start
    int energy_mode;
    int defer_weight;           // sharded runs with energy_mode 0: the step
                                // kernel leaves the branching weight to
                                // multi_weight_kernel (the stale-slot energy
                                // is indexed by GLOBAL position, known once
                                // the counts of all ranks are)
};

// Sharded runs (one rank per GPU).  The reference's per-slot array behind
// quirk Q1 (qmc_base/jastrow/dmc.py:810,936) is indexed by the position of
// a walker in the ONE ordered ensemble; here every rank keeps a replica of
// that global array and the ranks exchange, once per step, the slab of
// values each of them writes (an all-gather of 8 bytes per slot that runs
// next to the step kernel).
struct DmcMulti {
    double *aglob;          // [R * cap]  A[global position]
    double *slab;           // [cap]      E_prev[parent] of my children
    double *gathered;       // [R * cap]  the slabs of all ranks
    double *vred;           // [R + 2]    {sum E, W, W_0 .. W_{R-1}}
    long long *offs;        // [R + 1]    first global position of rank q
    int R, rank, cap;
};

struct DmcLog {
    double *energy, *weight, *ref_energy, *accum_energy;
    unsigned long long *num_walkers;
};

constexpr int BR_THREADS = 256;
constexpr int BR_ITEMS = 4;
constexpr int BR_TILE = BR_THREADS * BR_ITEMS;

__device__ __forceinline__ long long block_excl_scan(long long v,
                                                     long long *total)
{
    // exclusive scan of one value per thread over a BR_THREADS CTA
    __shared__ long long wsum[BR_THREADS / 32];
    __shared__ long long tot;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        long long w = (lane < BR_THREADS / 32) ? wsum[lane] : 0;
        long long winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long n = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += n;
        }
        if (lane < BR_THREADS / 32) wsum[lane] = winc - w;
        if (lane == 31) tot = winc;
    }
    __syncthreads();
    long long r = inc - v + wsum[wid];
    if (total) *total = tot;
    __syncthreads();
    return r;
}

// Sum of one value per thread over a BR_THREADS CTA, on every thread: warp
// shuffles, then the warp totals.  The tree has a fixed shape, so the
// floating-point result does not depend on scheduling.
template <typename T>
__device__ __forceinline__ T block_sum(T v)
{
    __shared__ T part[BR_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) part[wid] = v;
    __syncthreads();
    T r = part[0];
#pragma unroll
    for (int w = 1; w < BR_THREADS / 32; ++w) r += part[w];
    __syncthreads();            // part may be written again
    return r;
}

// Programmatic dependent launch (sm_90+): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may become resident
// while its predecessor in the stream is still running.  Every such kernel
// first waits here for the predecessor's COMPLETION (all its memory
// operations visible), then lets its own successor in: the launch latency
// and the CTA scheduling of the three kernels of a DMC time step overlap
// the tail of the kernel before.  Without the launch attribute both calls
// return at once.
__device__ __forceinline__ void pdl_enter()
{
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
}

// True in exactly one CTA of the grid: the last one to get here.  Its reads
// (through __ldcg) see everything the other CTAs wrote before their call.
__device__ __forceinline__ bool last_cta_done(unsigned int *counter)
{
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int n = atomicAdd(counter, 1u);
        is_last = (n == gridDim.x - 1);
        if (is_last) *counter = 0;          // ready for the next step
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// K4a: clone counts c_s = int(w_s + u_s) (qmc_base/dmc.py:641-643) and their
// per-CTA totals.
__global__ void __launch_bounds__(BR_THREADS)
branch_count_kernel(DmcBufs B, DmcConsts C, DmcMulti X, int fuse_weight)
{
    pdl_enter();
    const DmcCtl *ctl = B.ctl;
    const int Wp = ctl->W_prev;
    const int par = (int) (ctl->step & 1);
    const uint32_t step = (uint32_t) ctl->step;
    const long long pos0 = ctl->pos_base[par];
    double *w = B.weight[par];
    // Sharded runs (fuse_weight): the weights of the walkers about to branch
    // were left to this kernel by the previous step (DmcConsts::defer_weight);
    // they come from the global per-position stale-energy array exactly as in
    // multi_weight_kernel, with the E_ref that drove that step.
    const double eref_prev = ctl->eref[par ^ 1];
    const double *e_new = B.energy[par];
    const long long asize = (long long) X.R * X.cap;
    long long local = 0;
    int base = blockIdx.x * BR_TILE + threadIdx.x * BR_ITEMS;
#pragma unroll
    for (int i = 0; i < BR_ITEMS; ++i) {
        int s = base + i;
        int c = 0;
        if (s < Wp) {
            double u0, u1;
            rng_uniform2(C.seed, (uint32_t) (pos0 + s), 0u, step,
                         STREAM_BRANCH, u0, u1);
            double ws;
            if (fuse_weight) {
                const long long p = pos0 + s;
                const double e_old = p < asize ? X.aglob[p] : 0.0;
                ws = exp(-C.dt * ((e_new[s] + e_old) / 2 - eref_prev));
                w[s] = ws;
            } else {
                ws = w[s];
            }
            double x = ws + u0;
            c = (x >= (double) B.cap) ? B.cap : (int) x;
            if (c < 0) c = 0;
        }
        if (s < B.cap) B.cnt[s] = c;
        local += c;
    }
    const long long tot = block_sum(local);
    if (threadIdx.x == 0) B.blocksum[blockIdx.x] = tot;
}

// K7: population control (qmc_base/dmc.py:758-771) from the (global) sums in
// ctl->red, the per-step log, and the hand-over to the next step.  Runs
// before or NEXT TO the step kernel of the same time step (inline at the end
// of branch_fill_kernel on one rank, on a second stream after the all-reduce
// on several), which therefore takes its step index from ctl->tcur and only
// reads what this function does not write: W, and E_ref of parity t & 1.
__device__ __forceinline__ void dmc_finalize(const DmcBufs &B,
                                             const DmcConsts &C,
                                             const DmcLog &L)
{
    DmcCtl *ctl = B.ctl;
    double sE = ctl->red[0], sW = ctl->red[1];
    ctl->tot_e += sE;
    ctl->tot_w += sW;
    double accum = ctl->tot_e / ctl->tot_w;
    double eref = accum - C.nwc_over_dt * log(sW / C.target);
    long long t = ctl->step;
    ctl->eref[(t + 1) & 1] = eref;
    ctl->last_energy = sE;
    ctl->last_weight = sW;
    ctl->last_accum = accum;
    ctl->W_global = sW;
    long long i = t - ctl->block_step0;
    if (L.energy) {
        L.energy[i] = sE;
        L.weight[i] = sW;
        L.num_walkers[i] = (unsigned long long) (sW + 0.5);
        L.ref_energy[i] = eref;
        L.accum_energy[i] = accum;
    }
    ctl->W_prev = ctl->W;
    ctl->step = t + 1;
}

// K4c (+K7): children occupy consecutive slots in parent order, truncated at
// the capacity (qmc_base/dmc.py:644-653); per-CTA partial of sum E over the
// cloned parents (= state_energy, qmc_base/dmc.py:759-760).  The last CTA to
// finish adds the partials in a fixed order into ctl->red and, on a single
// rank (finalize != 0), runs the population control; with several ranks the
// host all-reduces ctl->red first and then launches dmc_finalize_kernel.
__global__ void __launch_bounds__(BR_THREADS)
branch_fill_kernel(DmcBufs B, DmcConsts C, DmcLog L, int finalize)
{
    pdl_enter();
    const DmcCtl *ctl = B.ctl;
    const int par = (int) (ctl->step & 1);
    const double *e = B.energy[par];
    int base = blockIdx.x * BR_TILE + threadIdx.x * BR_ITEMS;
    int c[BR_ITEMS];
    long long local = 0;
#pragma unroll
    for (int i = 0; i < BR_ITEMS; ++i) {
        int s = base + i;
        c[i] = (s < B.cap) ? B.cnt[s] : 0;
        local += c[i];
    }
    // children of the CTAs before this one: every CTA adds up the per-CTA
    // totals itself (a few hundred values at most; saves the scan pass and
    // its grid-wide hand-over in branch_count_kernel)
    long long before = 0;
    for (int b = threadIdx.x; b < (int) blockIdx.x; b += BR_THREADS)
        before += B.blocksum[b];
    const long long boff = block_sum(before);
    long long off = block_excl_scan(local, nullptr) + boff;
    double esum = 0.0;
#pragma unroll
    for (int i = 0; i < BR_ITEMS; ++i) {
        int s = base + i;
        int placed = 0;
        for (int k = 0; k < c[i]; ++k) {
            long long slot = off + k;
            if (slot >= B.cap) break;
            B.ref[slot] = s;
            ++placed;
        }
        if (placed) esum += (double) placed * e[s];
        off += c[i];
    }
    // fixed-shape CTA reduction: deterministic
    const double epart = block_sum(esum);
    if (threadIdx.x == 0) B.epart[blockIdx.x] = epart;
    if (!last_cta_done(&B.ctl->done_fill)) return;
    // W of this step: all children, truncated at the capacity
    long long kids = 0;
    double acc = 0.0;
    for (int b = threadIdx.x; b < B.nblk; b += BR_THREADS) {
        kids += B.blocksum[b];
        acc += __ldcg(B.epart + b);
    }
    long long total = block_sum(kids);
    const double etot = block_sum(acc);
    if (threadIdx.x == 0) {
        DmcCtl *c = B.ctl;
        c->total_children = total;
        if (total > B.cap) { c->capacity_hits += 1; total = B.cap; }
        c->W = (int) total;
        B.ctl->red[0] = etot;
        B.ctl->red[1] = (double) B.ctl->W;
        B.ctl->tcur = B.ctl->step;
        if (finalize) dmc_finalize(B, C, L);
    }
}

// First launch of every block: the per-step log of the block starts at the
// current step.  Keeping this on the device makes a block a launch sequence
// whose arguments never change, i.e. one CUDA graph replayed per block.
__global__ void dmc_block_begin_kernel(DmcBufs B)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
        B.ctl->block_step0 = B.ctl->step;
}

// On-the-fly reblocking of the five per-step series of a block
// (stats/reblock.py:525-604 `_on_the_fly_obj_create`, accumulated over the
// blocks like `on_the_fly_obj_data_update` :927-948): for every order k the
// sum and the sum of squares of the means of blocks of 2^k consecutive steps,
// built pairwise from the two means of order k-1 exactly as the reference
// does, so the tables are bit-identical to reblocking the shipped series.
// Thread c = series c (energy, weight, num_walkers, ref_energy, accum_energy).
constexpr int RB_COLS = 5;
constexpr int RB_MAX_ORDERS = 40;

struct ReblockTables {
    double *sum, *sqr;              // [RB_COLS][K]
    long long *nblk;                // [RB_COLS][K]
    int K;                          // orders kept: 0 .. K-1
};

__global__ void reblock_series_kernel(DmcLog L, long long nts,
                                      int block_max_order, ReblockTables T)
{
    const int c = threadIdx.x;
    if (blockIdx.x != 0 || c >= RB_COLS) return;
    const int kmax = block_max_order < T.K - 1 ? block_max_order : T.K - 1;
    // the table of THIS block first (on_the_fly_obj_create), added to the
    // running table afterwards (on_the_fly_obj_data_update): the same
    // association of the sums as the reference's
    double means[RB_MAX_ORDERS][2], sum[RB_MAX_ORDERS], sqr[RB_MAX_ORDERS];
    long long nb[RB_MAX_ORDERS];
    for (int k = 0; k <= kmax; ++k) { sum[k] = 0.0; sqr[k] = 0.0; nb[k] = 0; }
    for (long long idx = 0; idx < nts; ++idx) {
        double v;
        switch (c) {
        case 0: v = L.energy[idx]; break;
        case 1: v = L.weight[idx]; break;
        case 2: v = (double) L.num_walkers[idx]; break;
        case 3: v = L.ref_energy[idx]; break;
        default: v = L.accum_energy[idx]; break;
        }
        means[0][idx & 1] = v;
        sum[0] = __dadd_rn(sum[0], v);
        sqr[0] = __dadd_rn(sqr[0], __dmul_rn(v, v));
        nb[0] += 1;
        long long bs = 1;
        for (int k = 1; k <= kmax; ++k) {
            bs <<= 1;
            if ((idx + 1) % bs) break;
            const long long bidx = (idx + 1) / bs - 1;
            const double m = __ddiv_rn(
                __dadd_rn(means[k - 1][0], means[k - 1][1]), 2.0);
            means[k][bidx & 1] = m;
            sum[k] = __dadd_rn(sum[k], m);
            sqr[k] = __dadd_rn(sqr[k], __dmul_rn(m, m));
            nb[k] += 1;
        }
    }
    for (int k = 0; k <= kmax; ++k) {
        T.sum[c * T.K + k] = __dadd_rn(T.sum[c * T.K + k], sum[k]);
        T.sqr[c * T.K + k] = __dadd_rn(T.sqr[c * T.K + k], sqr[k]);
        T.nblk[c * T.K + k] += nb[k];
    }
}

// K7 on several ranks: after the all-reduce of ctl->red.
__global__ void dmc_finalize_kernel(DmcBufs B, DmcConsts C, DmcLog L)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    dmc_finalize(B, C, L);
}

// ---------------------------------------------------------------------------
// K3: the fused DMC step.  For every live child slot s with parent r:
//   gather (z, F) of r -> drift-diffusion move + recast -> tables ->
//   pair drift / local energy -> branching weight -> write the child.
// Reference: evolve_state_inner / evolve_system / ith_diffusion
// (qmc_base/jastrow/dmc.py:634-673, 743-951), recast (mrbp_qmc/dmc.py:453).
// ---------------------------------------------------------------------------
#ifndef QMCB_STEP_THREADS
#define QMCB_STEP_THREADS 256
#endif
#ifndef QMCB_STEP_MINCTAS
#define QMCB_STEP_MINCTAS (QMCB_TB == 2 ? 3 : 2)
#endif
// FAST: the per-particle transcendentals come from the model's node tables
// (TrigTab); needs positions in [0, L], i.e. the recast interval of the
// sampling equal to the supercell (checked by the host).
template <bool FAST>
__global__ void __launch_bounds__(QMCB_STEP_THREADS, QMCB_STEP_MINCTAS)
dmc_step_kernel(const __grid_constant__ DevModel M, GroupGeom geom, DmcBufs B,
                DmcConsts C)
{
    pdl_enter();
    const DmcCtl *ctl = B.ctl;
    const int W = ctl->W;
    const long long s0 = (long long) blockIdx.x * geom.G;
    if (s0 >= W) return;
    const long long t = ctl->tcur;      // not ctl->step: the population control
                                        // of this step may still be in flight
    const int par = (int) (t & 1);
    const double eref = ctl->eref[par];
    const double *pconfs = B.confs[par];
    const double *penergy = B.energy[par];
    double *nconfs = B.confs[par ^ 1];
    double *nenergy = B.energy[par ^ 1];
    double *nweight = B.weight[par ^ 1];

    GroupSmem sm = group_smem(geom);
    GroupIdx x = group_index(M, geom);
    const int N = M.nop;
    const bool vec_ok = (N % 2) == 0;
    const long long s = s0 + x.g;
    const bool active = x.in_group && s < W;
    const int nvalid = min(TB, N - TB * x.I);

    double z[TB] = {};
    int r = 0;
    if (active) {
        r = B.ref[s];
        double zp[TB], fp[TB];
        const double *pc = pconfs + (long long) r * 2 * N;
        load4(pc, x.I, nvalid, vec_ok, zp);
        load4(pc + N, x.I, nvalid, vec_ok, fp);
        double nrm[TB];
        // the normals of a child are keyed by its PARENT's global position
        // and by which of the parent's copies it is: both are known without
        // this step's collective, and do not depend on how the ensemble is
        // cut into ranks
        int clone = 0;
        while (clone < 0xffff && s - clone - 1 >= 0
               && B.ref[s - clone - 1] == r)
            ++clone;
        const uint32_t gp = (uint32_t) (ctl->pos_base[par] + r);
        {
            // one Philox call serves the four particles of a 4-block
            double n4[4];
            const int q4 = (TB * x.I) / 4, o4 = (TB * x.I) % 4;
            rng_normal4<FAST>(M.tt, C.seed, gp,
                              (uint32_t) q4 | ((uint32_t) clone << 16),
                              (uint32_t) t, STREAM_DIFFUSE, n4);
#pragma unroll
            for (int c = 0; c < TB; ++c) nrm[c] = n4[(o4 + c) & 3];
        }
#pragma unroll
        for (int c = 0; c < TB; ++c) {
            double zn = zp[c] + 2.0 * fp[c] * C.dt + C.sigma * nrm[c];
            z[c] = recast(zn, C.z_min, C.size);
        }
        // the new positions are final: write them now, not after the eval
        store4(nconfs + s * 2 * N, x.I, nvalid, vec_ok, z);
    }
    EvalOut o;
    group_eval<false, true, false, FAST>(M, sm, x.g, x.I, active, z, nvalid,
                                         o);
    if (active) {
        double *nc = nconfs + s * 2 * N;
        store4(nc + N, x.I, nvalid, vec_ok, o.F);
        if (x.I == 0) {
            nenergy[s] = o.energy;
            if (!C.defer_weight) {
                double e_parent = penergy[r];
                double e_old = (C.energy_mode == 0) ? B.slot_energy[s]
                                                    : e_parent;
                double mean = (o.energy + e_old) / 2;
                nweight[s] = exp(-C.dt * (mean - eref));
                B.slot_energy[s] = e_parent;
            }
        }
    }
}

// --- sharded runs -----------------------------------------------------------
// After branch_fill_kernel (finalize = 0): this rank's contribution to the
// per-step collectives.
__global__ void multi_pack_kernel(DmcBufs B, DmcMulti X, int with_slab)
{
    const DmcCtl *ctl = B.ctl;
    const int W = ctl->W;
    const int par = (int) (ctl->step & 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (with_slab && i < X.cap)
        X.slab[i] = (i < W) ? B.energy[par][B.ref[i]] : 0.0;
    if (i == 0) {
        X.vred[0] = ctl->red[0];
        X.vred[1] = ctl->red[1];
        for (int q = 0; q < X.R; ++q)
            X.vred[2 + q] = (q == X.rank) ? (double) W : 0.0;
    }
}

// After the all-reduce of vred: global sums -> population control, and the
// global positions of every rank's children.
__global__ void multi_finalize_kernel(DmcBufs B, DmcConsts C, DmcLog L,
                                      DmcMulti X)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    DmcCtl *ctl = B.ctl;
    ctl->red[0] = X.vred[0];
    ctl->red[1] = X.vred[1];
    long long o = 0;
    for (int q = 0; q < X.R; ++q) {
        X.offs[q] = o;
        o += (long long) (X.vred[2 + q] + 0.5);
    }
    X.offs[X.R] = o;
    const long long t = ctl->step;
    // this step's children are the parents of the next one
    ctl->pos_base[(t + 1) & 1] = X.offs[X.rank];
    dmc_finalize(B, C, L);
}

// Branching weight of this rank's children from the GLOBAL stale-slot array
// (qmc_base/jastrow/dmc.py:810,820-821).  Runs after the step kernel.
__global__ void multi_weight_kernel(DmcBufs B, DmcConsts C, DmcMulti X)
{
    const DmcCtl *ctl = B.ctl;
    const int W = ctl->W;
    const long long t = ctl->tcur;
    const int par = (int) (t & 1);
    const double eref = ctl->eref[par];
    const long long base = X.offs[X.rank];
    const long long asize = (long long) X.R * X.cap;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < W;
         s += gridDim.x * blockDim.x) {
        const long long p = base + s;
        const double e_old = p < asize ? X.aglob[p] : 0.0;
        const double mean = (B.energy[par ^ 1][s] + e_old) / 2;
        B.weight[par ^ 1][s] = exp(-C.dt * (mean - eref));
    }
}

// A[position] <- E_prev[parent] for the children of every rank
// (qmc_base/jastrow/dmc.py:936); positions beyond the population keep their
// values, like the reference's dead slots.
__global__ void multi_apply_kernel(DmcMulti X)
{
    const long long total = (long long) X.R * X.cap;
    for (long long idx = blockIdx.x * (long long) blockDim.x + threadIdx.x;
         idx < total; idx += (long long) gridDim.x * blockDim.x) {
        const int q = (int) (idx / X.cap);
        const long long i = idx - (long long) q * X.cap;
        const long long o = X.offs[q];
        if (i < X.offs[q + 1] - o && o + i < total)
            X.aglob[o + i] = X.gathered[idx];
    }
}

// The part of the global array beyond the live population, which the last
// rank's local array carries on import (restart / initial state).
__global__ void multi_tail_kernel(DmcMulti X, long long n_last)
{
    const long long total = (long long) X.R * X.cap;
    const long long T = X.offs[X.R];
    const double *src = X.gathered + (long long) (X.R - 1) * X.cap;
    for (long long i = n_last + blockIdx.x * (long long) blockDim.x
                       + threadIdx.x;
         i < X.cap; i += (long long) gridDim.x * blockDim.x) {
        const long long p = T + (i - n_last);
        if (p < total) X.aglob[p] = src[i];
    }
}

// Local view of the global array for export: out[i] = A[base + i].
__global__ void multi_export_kernel(DmcMulti X, long long base, double *out)
{
    const long long total = (long long) X.R * X.cap;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
         i < X.cap; i += (long long) gridDim.x * blockDim.x)
        out[i] = base + i < total ? X.aglob[base + i] : 0.0;
}

// Gather of the yielded ("actual") population: slot s <- parent ref[s].
__global__ void gather_state_kernel(const double *pconfs,
                                    const double *penergy, const int *ref,
                                    int W, int cap, int N, bool identity,
                                    double *oconfs, double *oenergy,
                                    double *oweight, unsigned char *omask,
                                    long long *oref)
{
    long long total = (long long) cap * 2 * N;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
         i < total; i += (long long) gridDim.x * blockDim.x) {
        long long s = i / (2 * N);
        int k = (int) (i - s * 2 * N);
        double v = 0.0;
        if (s < W) {
            int r = identity ? (int) s : ref[s];
            v = pconfs[(long long) r * 2 * N + k];
        }
        if (oconfs) oconfs[i] = v;
        if (k == 0) {
            int r = (s < W) ? (identity ? (int) s : ref[s]) : 0;
            if (oenergy) oenergy[s] = (s < W) ? penergy[r] : 0.0;
            if (oweight) oweight[s] = (s < W) ? 1.0 : 0.0;
            if (omask) omask[s] = (s < W) ? 0 : 1;
            if (oref) oref[s] = (s < W) ? r : 0;
        }
    }
}

// ---------------------------------------------------------------------------
// Roofline denominator: sustained DFMA rate (8 independent chains/thread).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b);
            x3 = fma(x3, a, b); x4 = fma(x4, a, b); x5 = fma(x5, a, b);
            x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace qmcb
