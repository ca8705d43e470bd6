/*
 * qmcb200.h -- C ABI of libqmcb200.so, the B200 (sm_100a) walker-ensemble
 * engine for the mrbp_qmc (multi-rods, Bijl-Jastrow) Bose-gas model.
 *
 * The reference (oarodriguez/PhD-QMCLib v0.17.0) is pure Python + Numba and has
 * no FFI of its own; each entry point below names the reference function(s)
 * it replaces.  Paths are relative to src/phd_qmclib/ of the reference.
 *
 * Conventions
 *   - every array is C-contiguous; float64 unless stated; the CALLER owns
 *     every host buffer, the library never keeps a host pointer after return;
 *   - a configuration is (2, N): row 0 positions, row 1 drift
 *     (qmc_base/jastrow/model.py:30-38);
 *   - functions return 0 on success, a negative qmcb_status otherwise; the
 *     message is available through qmcb_last_error();
 *   - a handle is bound to one CUDA device and is not re-entrant (one host
 *     thread at a time), like the reference's generators.
 */
#ifndef QMCB200_H
#define QMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMCB_API __attribute__((visibility("default")))

typedef enum {
    QMCB_OK = 0,
    QMCB_ERR_INVALID = -1,      /* bad argument / unsupported size        */
    QMCB_ERR_CUDA = -2,         /* CUDA runtime error (sticky)            */
    QMCB_ERR_STATE = -3,        /* call order (e.g. run before init)      */
    QMCB_ERR_NCCL = -4          /* NCCL error or NCCL not loadable        */
} qmcb_status;

typedef struct qmcb_handle qmcb_handle;

/*
 * The 25 model scalars, in the order the reference ships them through its own
 * params -> float64-array transforms (mrbp_qmc/model.py:571-686):
 *   model[12] = Params     (mrbp_qmc/model.py:40-54)
 *               lattice_depth, lattice_ratio, interaction_strength,
 *               boson_number, supercell_size, tbf_contact_cutoff,
 *               defect_magnitude, defects_sep, well_width, barrier_width,
 *               is_free, is_ideal
 *   obf[7]    = OBFParams  (:57-65)  lattice_depth, lattice_ratio, well_width,
 *               barrier_width, param_e0, param_k1, param_kp1
 *   tbf[6]    = TBFParams  (:68-75)  supercell_size, tbf_contact_cutoff,
 *               param_k2, param_beta, param_r_off, param_am
 */
typedef struct {
    double model[12];
    double obf[7];
    double tbf[6];
} qmcb_model_params;

/* DMC sampling parameters: mrbp_qmc/dmc.py:144-250 (Sampling, ddf_params,
 * density_params, ssf_params). */
typedef struct {
    double time_step;               /* Sampling.time_step                     */
    double nwc_factor;              /* num_walkers_control_factor             */
    double lower_bound;             /* DDFParams.lower_bound (= 0)            */
    double upper_bound;             /* DDFParams.upper_bound (= L)            */
    int64_t max_num_walkers;        /* GLOBAL capacity (all ranks)            */
    int64_t target_num_walkers;     /* GLOBAL target                          */
    uint64_t rng_seed;
    int32_t energy_mode;            /* 0: reference (stale slot energy, quirk
                                       Q1, jastrow/dmc.py:810); 1: parent's   */
    int32_t ssf_num_modes;          /* 0 = S(k) estimator off                 */
    int32_t ssf_as_pure;
    int32_t density_num_bins;       /* 0 = density estimator off              */
    int32_t density_as_pure;
    int32_t reserved0;
    int64_t ssf_pfw_nts;            /* SSFEstSpec.pfw_num_time_steps          */
    int64_t density_pfw_nts;
    int64_t local_capacity;         /* slots on THIS rank; 0 = max_num_walkers
                                       (single GPU)                           */
} qmcb_dmc_params;

/* VMC sampling parameters: mrbp_qmc/vmc.py:71-128 (Sampling, tpf_params,
 * ssf_params).  The engine advances `num_chains` independent chains; chain c
 * is the reference's single-chain algorithm (qmc_base/vmc.py:557-648). */
typedef struct {
    double move_spread;             /* proposal 0: width of the uniform move;
                                       proposal 1: sigma = sqrt(time_step)    */
    double lower_bound;
    double upper_bound;
    uint64_t rng_seed;
    int64_t chain_offset;           /* global index of local chain 0          */
    int32_t ssf_num_modes;          /* 0 = off                                */
    int32_t proposal;               /* 0: uniform  (qmc_base/vmc.py:401-415);
                                       1: gaussian (qmc_base/vmc_ndf.py:44-62,
                                          mrbp_qmc/vmc_ndf.py:24-60)          */
} qmcb_vmc_params;

/* Scalars of a DMC State (qmc_base/dmc.py:117-127). */
typedef struct {
    double energy;                  /* sum of E_L over the live walkers       */
    double weight;
    double ref_energy;
    double accum_energy;
    double total_energy;            /* running totals of states_generator     */
    double total_weight;            /*   (qmc_base/dmc.py:733-765)            */
    int64_t num_walkers;            /* live walkers on this rank              */
    int64_t max_num_walkers;        /* local capacity                         */
    int64_t step;                   /* time steps done since init             */
    int64_t capacity_hits;          /* steps in which branching was truncated
                                       (quirk Q8; the reference is silent)    */
} qmcb_state_scalars;

/* ---- lifetime ---------------------------------------------------------- */
QMCB_API int qmcb_create(const qmcb_model_params *params, int device,
                         qmcb_handle **out);
QMCB_API void qmcb_destroy(qmcb_handle *h);
/* Message of the last failure on this handle (or of the last failed
 * qmcb_create when h == NULL).  Never NULL. */
QMCB_API const char *qmcb_last_error(const qmcb_handle *h);
QMCB_API const char *qmcb_version(void);

/* ---- fixed-configuration evaluation ------------------------------------ */
/* Replaces core_funcs.wf_abs_log (qmc_base/jastrow/model.py:336-368),
 * .energy (:748-775), .drift (:528-566) on a batch of configurations.
 * confs [nconf][2][N] (row 1 ignored).  Any of lnpsi[nconf], energy[nconf],
 * drift[nconf][N] may be NULL. */
QMCB_API int qmcb_model_eval(qmcb_handle *h, const double *confs,
                             int64_t nconf, double *lnpsi, double *energy,
                             double *drift);
/* Same, all pointers in DEVICE memory of the handle's device. */
QMCB_API int qmcb_model_eval_device(qmcb_handle *h, const double *d_confs,
                                    int64_t nconf, double *d_lnpsi,
                                    double *d_energy, double *d_drift);

/* Replaces core_funcs.fourier_density (qmc_base/jastrow/model.py:968-1004)
 * for k_m = m 2 pi / L, m = 0..num_modes-1 (mrbp_qmc/dmc.py:633), with the
 * slots of SSFPartSlot (qmc_base/dmc.py:75-85): out [nconf][num_modes][3] =
 * (|rho_k|^2, Re rho_k, Im rho_k). */
QMCB_API int qmcb_fourier_density(qmcb_handle *h, const double *confs,
                                  int64_t nconf, int32_t num_modes,
                                  double *out);

/* Replaces core_funcs.one_body_density (qmc_base/jastrow/model.py:934-965,
 * per-particle term :859-931) and its broadcasting gufunc
 * PhysicalFuncs.one_body_density (:1069-1091):
 * out [nconf][num_offsets] = (1/N) sum_i Psi(z_i + offsets[s]) / Psi for each
 * configuration.  Offsets may lie outside the box (the wave function is
 * periodic in L). */
QMCB_API int qmcb_one_body_density(qmcb_handle *h, const double *confs,
                                   int64_t nconf, const double *offsets,
                                   int32_t num_offsets, double *out);
/* Same, all pointers in DEVICE memory; asynchronous on the engine's stream. */
QMCB_API int qmcb_one_body_density_device(qmcb_handle *h,
                                          const double *d_confs,
                                          int64_t nconf,
                                          const double *d_offsets,
                                          int32_t num_offsets,
                                          double *d_out);

/* Replaces the gufunc PhysicalFuncs.fourier_density
 * (qmc_base/jastrow/model.py:1093-1122) for an ARBITRARY momentum set:
 * out [nconf][nk][2] = (Re, Im) of rho_k = sum_i exp(i kz z_i). */
QMCB_API int qmcb_fourier_density_k(qmcb_handle *h, const double *confs,
                                    int64_t nconf, const double *kz,
                                    int32_t nk, double *out);

/* ---- trial-wave-function optimisation (correlated sampling) ------------- */
/* Swap the model scalars of a live handle (CSWFOptimizer.update_spec,
 * mrbp_qmc/model.py:852-862, changes tbf_contact_cutoff and with it the six
 * two-body scalars).  boson_number must not change. */
QMCB_API int qmcb_set_model_params(qmcb_handle *h,
                                   const qmcb_model_params *params);
/* Keep the configuration set of the optimiser on the device
 * (CSWFOptimizer.sys_conf_set / ini_wf_abs_log_set, mrbp_qmc/model.py:828-832):
 * confs [nconf][2][N], ini_lnpsi [nconf] (NULL: evaluate it with the handle's
 * current parameters). */
QMCB_API int qmcb_cs_load(qmcb_handle *h, const double *confs, int64_t nconf,
                          const double *ini_lnpsi);
/* Replaces CSWFOptimizer.principal_function
 * (qmc_base/jastrow/model.py:1186-1206): ln|Psi| and E_L of the loaded set
 * under the `trial` scalars (NULL: the handle's own; wf_abs_log_and_energy_set,
 * mrbp_qmc/model.py:889-903), weights exp(2 (ln|Psi| - ln|Psi_0|)), and the
 * weighted variance of E_L (weighed_variance, :1147-1165), all on the device.
 * ref_energy = the weighted mean energy; lnpsi/energy [nconf] host, any
 * output may be NULL. */
QMCB_API int qmcb_cs_variance(qmcb_handle *h, const qmcb_model_params *trial,
                              double *variance, double *ref_energy,
                              double *lnpsi, double *energy);

/* ---- DMC ---------------------------------------------------------------- */
/* Replaces Sampling.build_state (mrbp_qmc/dmc.py:268-328) +
 * prepare_state_data (qmc_base/jastrow/dmc.py:1030-1174) + the three-buffer
 * set-up of states_generator (qmc_base/dmc.py:700-735).
 * ini_confs [n][2][N] host; ref_energy = NaN -> mean local energy.
 * global_slot_offset: global index of this rank's slot 0 (RNG keying). */
QMCB_API int qmcb_dmc_init(qmcb_handle *h, const qmcb_dmc_params *params,
                           const double *ini_confs, int64_t n,
                           double ref_energy, int64_t global_slot_offset);

/* Restart from a saved State (ProcInput.from_result,
 * mrbp_qmc/dmc_exec/proc.py:131-143; layout of qmc_exec/dmc/io.py:35-57).
 * confs [num_walkers][2][N], energy/weight [num_walkers] are the walkers of
 * state.confs/props with mask == False; slot_energy [capacity] (may be NULL:
 * then = energy) restores the persistent per-slot array behind quirk Q1. */
QMCB_API int qmcb_dmc_set_state(qmcb_handle *h, const qmcb_dmc_params *params,
                                const double *confs, const double *energy,
                                const double *weight,
                                const double *slot_energy,
                                const qmcb_state_scalars *scalars,
                                int64_t global_slot_offset);

/* Replaces one `next()` of CoreFuncs.blocks (qmc_base/dmc.py:815-971):
 * nts iterations of states_generator (:664-787) = branching
 * (sync_branching_spec :614-655), evolve_state
 * (qmc_base/jastrow/dmc.py:830-951), population control (:758-771), and, iff
 * eval_estimators != 0, the density / S(k) estimators
 * (qmc_base/jastrow/dmc.py:195-302, 363-573).
 * Outputs (host, may each be NULL): energy/weight/ref_energy/accum_energy
 * [nts], num_walkers [nts] (uint64, as qmc_base/dmc.py:376), density
 * [nts][num_bins], ssf [nts][num_modes][3].  With several ranks the series
 * hold GLOBAL values; density / ssf hold this rank's partial sums (the caller
 * all-reduces them per block). */
QMCB_API int qmcb_dmc_run_block(qmcb_handle *h, int64_t nts,
                                int32_t eval_estimators, double *energy,
                                double *weight, uint64_t *num_walkers,
                                double *ref_energy, double *accum_energy,
                                double *density, double *ssf);

/* The yielded State (qmc_base/dmc.py:117-127) of the last step: the
 * post-branching, pre-move ("actual") population.  confs [capacity][2][N],
 * energy/weight [capacity], mask [capacity] (1 = dead slot),
 * cloning_ref [capacity]; any may be NULL. */
QMCB_API int qmcb_dmc_get_state(qmcb_handle *h, double *confs, double *energy,
                                double *weight, uint8_t *mask,
                                int64_t *cloning_ref,
                                qmcb_state_scalars *scalars);
/* The evolved population that the next step will branch from (the
 * reference's aux_next buffers) -- what a bit-exact restart needs. */
QMCB_API int qmcb_dmc_get_next(qmcb_handle *h, double *confs, double *energy,
                               double *weight, double *slot_energy,
                               qmcb_state_scalars *scalars);

/* On-the-fly reblocking of the per-step series on the device, so that long
 * runs need not ship the series themselves.  Replaces, for the five series of
 * PropsData (energy, weight, num_walkers, ref_energy, accum_energy; rows of
 * the tables in that order), stats.reblock._on_the_fly_obj_create
 * (stats/reblock.py:525-604) applied to each block's series and
 * on_the_fly_obj_data_update (:927-948) over the blocks run since the reset:
 * for order k = 0..max_order, the sum and the sum of squares of the means of
 * blocks of 2^k consecutive steps and their number (fields MEANS, MEANS_SQR,
 * NUM_BLOCKS of otf_data_dtype, :436-441; BLOCK_SIZE = 2^k).  A block of nts
 * steps feeds the orders up to floor(log2 nts).  max_order < 0 switches the
 * accumulators off.  Tables are [5][max_order + 1], host, each may be NULL. */
QMCB_API int qmcb_dmc_reblock_reset(qmcb_handle *h, int32_t max_order);
QMCB_API int qmcb_dmc_reblock_get(qmcb_handle *h, double *means_sum,
                                  double *means_sqr_sum, int64_t *num_blocks);

/* on != 0: bracket every step-kernel launch with CUDA events so that
 * qmcb_last_block_stats can report the step kernel's own device time. */
QMCB_API int qmcb_set_profiling(qmcb_handle *h, int32_t on);

/* Device time (ms, CUDA events on the engine's stream) and number of kernel
 * launches of the last qmcb_dmc_run_block / qmcb_vmc_run_block. */
QMCB_API int qmcb_last_block_stats(qmcb_handle *h, double *total_ms,
                                   double *step_kernel_ms, int64_t *launches);

/* Diagnostic: sustained DFMA throughput of `device` (8 independent chains
 * per thread, all SMs), in TFLOP/s -- the fp64 roofline denominator that
 * MEASURED_PEAKS.json lacks.  ms (may be NULL) = best kernel time. */
QMCB_API int qmcb_measure_fp64_peak(int device, double *tflops, double *ms);
/* Same kernel launched back to back for `seconds` of wall time: the rate the
 * FP64 pipe sustains under the power cap (denominator for a kernel timed
 * inside a long step). */
QMCB_API int qmcb_measure_fp64_sustained(int device, double seconds,
                                         double *tflops);

/* The engine's CUDA stream (a cudaStream_t), so that a caller can bracket
 * calls with its own events on the launching stream. */
QMCB_API void *qmcb_stream(qmcb_handle *h);
/* Page-locked host buffers for the host-pointer entry points (copies from
 * pageable memory are staged by the driver and are several times slower). */
QMCB_API void *qmcb_host_alloc(int64_t bytes);
QMCB_API void qmcb_host_free(void *p);

/* ---- multi-GPU (one process per GPU) ------------------------------------ */
/* 128-byte NCCL unique id, made on rank 0 and broadcast by the caller's own
 * plumbing (torch.distributed).  Call qmcb_comm_init BEFORE qmcb_dmc_init /
 * qmcb_dmc_set_state.  The ranks then hold ordered slabs of ONE ensemble
 * (rank r: the walkers after those of the lower ranks) and walk exactly the
 * ensemble a single rank would walk with the same seed:
 *  - the RNG is keyed by a walker's GLOBAL position in that order (the
 *    engine derives every rank's first position from the counts of all
 *    ranks; global_slot_offset is ignored);
 *  - the reference's per-slot stale-energy array (quirk Q1,
 *    qmc_base/jastrow/dmc.py:810,936) is kept per global position,
 *    replicated on every rank, updated each step by an all-gather of 8 bytes
 *    per slot; `slot_energy` of qmcb_dmc_get_next / qmcb_dmc_set_state is the
 *    rank's view of it (entries of its walkers; the last rank's entries
 *    beyond its population are the tail beyond the ensemble);
 *  - per step one all-reduce of world_size + 2 doubles {sum E, W, counts}
 *    drives the population control (qmc_base/dmc.py:758-771); both
 *    collectives run on a second stream next to the step kernel;
 *  - the per-step series and the estimator tables returned by
 *    qmcb_dmc_run_block hold GLOBAL values on every rank;
 *  - inside a block the slabs are evened out (qmcb_dmc_rebalance) whenever
 *    they drift apart by more than 1 % or near their capacity, unless a
 *    pure estimator carries per-slot state through the block.
 * The truncation at capacity (quirk Q8) acts per slab: give local_capacity
 * slack beyond max_num_walkers / world_size and watch capacity_hits. */
QMCB_API int qmcb_comm_unique_id(uint8_t id[128]);
QMCB_API int qmcb_comm_init(qmcb_handle *h, const uint8_t id[128],
                            int32_t world_size, int32_t rank);
/* Order-preserving neighbour rebalance: even out the live
 * populations across ranks by shifting tail/head walkers between ring
 * neighbours.  Collective call.  moved (may be NULL) = walkers sent. */
QMCB_API int qmcb_dmc_rebalance(qmcb_handle *h, int64_t *moved);

/* Exchange plan of qmcb_dmc_rebalance, host arithmetic only (no device, no
 * communicator): counts[world] = live walkers per rank.  For peer p,
 * send[2p], send[2p+1] = (offset in the caller's current slab, count) of the
 * walkers p takes over; recv[2p], recv[2p+1] = (offset in the caller's NEW
 * slab, count) of the walkers p hands over (p == rank: the part that stays).
 * new_count (may be NULL) = the caller's population afterwards. */
QMCB_API int qmcb_rebalance_plan(const int64_t *counts, int32_t world,
                                 int32_t rank, int64_t *send, int64_t *recv,
                                 int64_t *new_count);

/* ---- VMC ---------------------------------------------------------------- */
/* Replaces Sampling.build_state (mrbp_qmc/vmc.py:145-170) for num_chains
 * chains: confs [num_chains][2][N].  Calling it again with the same
 * num_chains and ssf_num_modes re-initialises the chains in the device
 * buffers of the previous call (no allocation); page-locked host buffers
 * make this and the result copies below run at PCIe rate. */
QMCB_API int qmcb_vmc_init(qmcb_handle *h, const qmcb_vmc_params *params,
                           const double *confs, int64_t num_chains);
/* Replaces one `next()` of vmc CoreFuncs.blocks (qmc_base/vmc.py:670-770)
 * per chain.  Outputs host, chain-major: lnpsi/energy [C][ns], move_stat
 * [C][ns] (uint8), ssf [C][ns][M][3], accept_rate [C]; each may be NULL.
 * sum_* (may be NULL): per-chain block sums of energy and energy^2 [C][2] and
 * of the ssf slots [C][M][3], accumulated on the device -- the cheap path
 * when the per-step series are not needed. */
QMCB_API int qmcb_vmc_run_block(qmcb_handle *h, int64_t ns, double *lnpsi,
                                double *energy, uint8_t *move_stat,
                                double *ssf, double *accept_rate,
                                double *sum_energy, double *sum_ssf);
/* The same chain with EVERY state kept (reference `as_chain` /
 * `state_data_blocks`, qmc_base/vmc.py:773-902: confs (ns, 2, N) per chain,
 * used e.g. to seed a DMC run): one launch for the ns steps, the
 * configurations recorded on the device and shipped once.
 * confs [C][ns][2][N] (row 0 positions, row 1 drift); the other outputs as in
 * qmcb_vmc_run_block, each may be NULL except confs. */
QMCB_API int qmcb_vmc_run_chain(qmcb_handle *h, int64_t ns, double *lnpsi,
                                double *energy, uint8_t *move_stat,
                                double *confs, double *accept_rate);

/* One-body density matrix estimator g1(s) of the CURRENT state of every chain
 * at the displacements `offsets` (reference hook `CoreFuncs.one_body_density`,
 * qmc_base/jastrow/vmc.py:267-301, over `ith_one_body_density`
 * jastrow/model.py:859-965), evaluated where the chains live: out [C][S]. */
QMCB_API int qmcb_vmc_one_body_density(qmcb_handle *h, const double *offsets,
                                       int32_t num_offsets, double *out);

/* Current configurations [C][2][N] and ln|Psi| [C] (last_state). */
QMCB_API int qmcb_vmc_get_state(qmcb_handle *h, double *confs, double *lnpsi);

#ifdef __cplusplus
}
#endif
#endif /* QMCB200_H */
