// C ABI of libqmcb200.so (see include/qmcb200.h).
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>       // types and prototypes only: the library is dlopen'ed

#include "../../include/qmcb200.h"
#include "qmcb_kernels.cuh"
#include "qmcb_estimators.cuh"
#include "qmcb_vmc.cuh"

using namespace qmcb;

namespace {

constexpr int CS_BLOCKS = 296;      // CTAs of the column-sum kernels (2 / SM)

std::string g_create_error;

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
};

}  // namespace

// NCCL entry points, resolved at run time so that libqmcb200.so loads (and its
// single-GPU paths run) on a box without NCCL.  Inside a torch process the
// soname resolves to the NCCL torch already mapped.
struct NcclApi {
    void *dl = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitRankConfig) CommInitRankConfig = nullptr;  // optional
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string err;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.dl ? &api : nullptr;
    tried = true;
    const char *env = getenv("QMCB_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.dl) break;
    }
    if (!api.dl) {
        api.err = "libnccl.so.2 not loadable (set QMCB_NCCL_LIB)";
        return nullptr;
    }
#define QMCB_NCCL_SYM(name)                                                  \
    api.name = (decltype(api.name)) dlsym(api.dl, "nccl" #name);             \
    if (!api.name) {                                                         \
        api.err = "nccl" #name " missing";                                   \
        api.dl = nullptr;                                                    \
        return nullptr;                                                      \
    }
    QMCB_NCCL_SYM(GetUniqueId) QMCB_NCCL_SYM(CommInitRank)
    QMCB_NCCL_SYM(CommDestroy) QMCB_NCCL_SYM(AllReduce)
    QMCB_NCCL_SYM(AllGather) QMCB_NCCL_SYM(Send) QMCB_NCCL_SYM(Recv)
    QMCB_NCCL_SYM(GroupStart) QMCB_NCCL_SYM(GroupEnd)
    QMCB_NCCL_SYM(GetErrorString)
#undef QMCB_NCCL_SYM
    api.CommInitRankConfig = (decltype(api.CommInitRankConfig)) dlsym(
        api.dl, "ncclCommInitRankConfig");
    return &api;
}

struct qmcb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    qmcb_model_params params{};
    DevModel M{};
    GroupGeom geom{};
    int sm_count = 0;
    int max_smem = 0;
    std::string err;

    // node tables of the per-particle transcendentals (TrigTab, device)
    double4 *d_trig_z = nullptr, *d_trig_c = nullptr;

    // scratch for host-pointer model evaluation
    double *d_scratch = nullptr;
    size_t scratch_bytes = 0;

    // DMC
    bool dmc_ready = false;
    qmcb_dmc_params dp{};
    DmcBufs B{};
    DmcConsts C{};
    DmcLog L{};
    long long log_cap = 0;
    long long step_host = 0;
    int n_ini = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> step_ev;
    bool profile_steps = false;
    double last_total_ms = 0.0, last_step_ms = 0.0;
    long long last_launches = 0;

    // a block without estimators on one rank = one CUDA graph of
    // 1 + 3 nts kernel nodes, rebuilt only when nts or the buffers change
    cudaGraphExec_t block_graph = nullptr;
    long long block_graph_nts = 0;

    // on-the-fly reblocking tables of the per-step series (device)
    ReblockTables rb{};

    // estimators (device)
    double *ssf_aux[2] = {nullptr, nullptr};    // [cap][M][3] ping-pong
    double *ssf_iter = nullptr;                 // [log_cap][M][3]
    double *den_hist[2] = {nullptr, nullptr};   // [cap][B] (2nd: mixed only)
    double *den_total = nullptr;                // [2][B]
    double *den_iter = nullptr;                 // [log_cap][B]
    double *est_partial = nullptr;              // [CS_BLOCKS][max(3M, B)]
    int *den_hi = nullptr;                      // highest slot count seen
    // pure density without per-slot histograms (density_list_kernel)
    bool den_lists_mode = false;
    unsigned short *den_lists = nullptr;        // [steps][cap][N] bin indices
    int *den_wrec = nullptr;                    // [steps] population recorded
    long long den_lists_steps = 0;
    long long est_log_cap = 0;

    // correlated-sampling set of the wave-function optimiser (device)
    double *cs_confs = nullptr;         // [cs_n][2][N]
    double *cs_ln0 = nullptr;           // [cs_n] ln|Psi| the set was drawn from
    double *cs_work = nullptr;          // [2 cs_n + 4]: ln|Psi|, E_L, results
    long long cs_n = 0;

    // VMC
    bool vmc_ready = false;
    bool vmc_fast = false;      // block kernel on the node tables
    qmcb_vmc_params vp{};
    VmcState V{};
    long long vmc_chains = 0, vmc_gstep = 0;
    int vmc_first = 1;
    double *vmc_sum_e = nullptr, *vmc_sum_ssf = nullptr, *vmc_acc = nullptr;

    // multi-GPU
    cudaStream_t pc_stream = nullptr;   // all-reduce + population control of
    cudaEvent_t ev_branched = nullptr;  // a step run here, next to the step
    cudaEvent_t ev_controlled = nullptr;    // kernel (which needs neither)
    cudaEvent_t ev_weighted = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    long long *d_counts = nullptr;      // [world] live walkers per rank
    DmcMulti X{};                       // global stale-slot array and the
    bool multi_ready = false;           // per-step collective buffers
    bool weights_pending = false;       // the last step's weights are still
                                        // to be formed (by the next branching
                                        // or by materialize_weights)
    long long rebalance_every = 32;     // in-block check period (steps)
    long long rebalanced_in_block = 0;
};

#define CUDA_TRY(h, expr)                                                    \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) {                                             \
            char _b[512];                                                    \
            snprintf(_b, sizeof _b, "%s failed at %s:%d: %s", #expr,         \
                     __FILE__, __LINE__, cudaGetErrorString(_e));            \
            (h)->err = _b;                                                   \
            return QMCB_ERR_CUDA;                                            \
        }                                                                    \
    } while (0)

#define NCCL_TRY(h, expr)                                                    \
    do {                                                                     \
        ncclResult_t _r = (expr);                                            \
        if (_r != ncclSuccess) {                                             \
            char _b[512];                                                    \
            snprintf(_b, sizeof _b, "%s failed at %s:%d: %s", #expr,         \
                     __FILE__, __LINE__, nccl_api()->GetErrorString(_r));    \
            (h)->err = _b;                                                   \
            return QMCB_ERR_NCCL;                                            \
        }                                                                    \
    } while (0)

#define FAIL(h, code, msg)                                                   \
    do {                                                                     \
        (h)->err = (msg);                                                    \
        return (code);                                                       \
    } while (0)

namespace {

// Host-side constants derived from the 25 reference scalars.
bool build_model(const qmcb_model_params &p, DevModel &M, std::string &err)
{
    const double *m = p.model, *o = p.obf, *t = p.tbf;
    M.tt = TrigTab{};       // tables belong to the handle's own parameters
    double nopd = m[3];
    if (!(nopd >= 1) || nopd != std::floor(nopd) || nopd > 4096) {
        err = "boson_number must be an integer in [1, 4096]";
        return false;
    }
    M.nop = (int) nopd;
    M.nb = (M.nop + TB - 1) / TB;
    M.kmax = M.nb / 2;
    M.is_free = m[10] != 0.0;
    M.is_ideal = m[11] != 0.0;
    M.defects_sep = (int) m[7];
    if (M.defects_sep < 1) M.defects_sep = 1;
    M.L = m[4];
    if (!(M.L > 0)) { err = "supercell_size must be positive"; return false; }
    M.inv_L = 1.0 / M.L;
    double v0 = o[0], r = o[1];
    M.za = 1 / (1 + r);     // mrbp_qmc/model.py:420 recomputes it from r
    M.zb = r / (1 + r);
    M.e0 = o[4]; M.k1 = o[5]; M.kp1 = o[6];
    M.v0 = v0;
    M.vdef = m[6];
    M.ln_cf = 0.0;
    if (!M.is_free) {
        double sh = std::sinh(0.5 * std::sqrt(v0 - M.e0) * M.zb);
        M.ln_cf = 0.5 * std::log(1 + v0 / M.e0 * sh * sh);
    }
    double rm = std::fabs(t[1]);
    double k2 = t[2], beta = t[3], r_off = t[4], am = t[5];
    M.k2 = k2; M.beta = beta;
    M.k1_over_pi = M.k1 / M_PI;
    M.k2_over_pi = k2 / M_PI;
    M.ln_am = std::log(std::fabs(am));
    double s_m = std::sin(M_PI * rm / M.L);
    if (rm >= 0.5 * M.L) s_m = 1.0;
    if (!M.is_ideal && !(k2 > 0 && beta > 0)) {
        err = "two-body parameters k2 and beta must be positive";
        return false;
    }
    // near-branch units (qmcb_dev.cuh): drift in -k2, kinetic in k2^2
    double gam = M.is_ideal ? 1.0 : (M_PI / M.L) * std::sqrt(beta) / k2;
    M.inv_gam = 1.0 / gam;
    M.mu = M.is_ideal ? 0.0 : -(M_PI / M.L) * beta / k2;
    M.s_m_scaled = s_m / gam;
    M.ln_gam = std::log(gam);
    M.drift_unit = M.is_ideal ? 1.0 : -k2;
    M.kin_unit = M.is_ideal ? 1.0 : k2 * k2;
    M.inv_drift_unit = 1.0 / M.drift_unit;
    M.half_inv_kin_unit = 0.5 / M.kin_unit;
    double psi0 = k2 * r_off, psi1 = k2 * r_off - k2 * M.L;
    M.cpsi[0] = std::cos(psi0); M.spsi[0] = std::sin(psi0);
    M.cpsi[1] = std::cos(psi1); M.spsi[1] = std::sin(psi1);
    return true;
}

// Node tables for the fast per-particle path (TrigTab in qmcb_dev.cuh).  The
// node counts keep every residual angle below TRIG_EPS_MAX; a model that
// would need more than 2^16 nodes keeps the exact sincospi / exp path
// (M.tt.zt == nullptr).  Entries are rounded from long double.
int build_trig_tables(qmcb_handle *h)
{
    DevModel &M = h->M;
    M.tt = TrigTab{};
    cudaFree(h->d_trig_z); cudaFree(h->d_trig_c);
    h->d_trig_z = h->d_trig_c = nullptr;
    if (getenv("QMCB_NO_TRIG_TABLES")) return QMCB_OK;
    auto pow2_at_least = [](double x) {
        int n = 256;
        while (n < x && n < (1 << 17)) n <<= 1;
        return n;
    };
    const double per_eps = 0.5 / TRIG_EPS_MAX;      // nodes per radian
    const double k2 = M.is_ideal ? 0.0 : M.k2;
    const double k1 = M.is_free ? 0.0 : M.k1, kp1 = M.is_free ? 0.0 : M.kp1;
    const int nz = pow2_at_least(per_eps * std::max(M_PI, k2 * M.L));
    const int nc = pow2_at_least(per_eps * std::max(k1, kp1));
    // sinh/cosh of the barrier argument must stay finite
    if (nz > (1 << 16) || nc > (1 << 16) || kp1 * 1.0 > 600.0)
        return QMCB_OK;
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<double4> zt(nz + 1), ct(nc + 1);
    for (int k = 0; k <= nz; ++k) {
        long double a = pi * k / nz;
        long double u = (long double) k2 * ((long double) M.L * k / nz);
        zt[k] = make_double4((double) sinl(a), (double) cosl(a),
                             (double) sinl(u), (double) cosl(u));
    }
    zt[nz].x = 0.0;     // sin(pi) exactly
    for (int j = 0; j <= nc; ++j) {
        long double zc = (long double) j / nc;
        long double w = (long double) k1 * (zc - 0.5L * M.za);
        long double y = (long double) kp1 * (zc - 1.0L + 0.5L * M.zb);
        ct[j] = make_double4((double) sinl(w), (double) cosl(w),
                             (double) sinhl(y), (double) coshl(y));
    }
    CUDA_TRY(h, cudaMalloc(&h->d_trig_z, zt.size() * sizeof(double4)));
    CUDA_TRY(h, cudaMalloc(&h->d_trig_c, ct.size() * sizeof(double4)));
    CUDA_TRY(h, cudaMemcpy(h->d_trig_z, zt.data(),
                           zt.size() * sizeof(double4),
                           cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_trig_c, ct.data(),
                           ct.size() * sizeof(double4),
                           cudaMemcpyHostToDevice));
    TrigTab &t = M.tt;
    t.zt = h->d_trig_z; t.ct = h->d_trig_c;
    t.nz = nz; t.nc = nc;
    t.z_scale = nz / M.L; t.eps_a = M_PI / nz; t.eps_u = k2 * M.L / nz;
    t.c_scale = nc; t.eps_w = k1 / nc; t.eps_b = kp1 / nc;
    return QMCB_OK;
}

// The step kernel may take the table path when the sampling recasts into
// exactly the supercell [0, L] the tables cover.
bool step_fast_ok(const qmcb_handle *h)
{
    return h->M.tt.zt != nullptr && h->C.z_min == 0.0 && h->C.size == h->M.L;
}

// CTA shape: threads per walker = nb; pack G walkers into a CTA so that few
// lanes idle, and pick the number of column-sum slots kept in shared memory
// (kc) so that shared memory does not cap the resident warps below what the
// registers allow.
bool choose_geom(const DevModel &M, int max_smem, GroupGeom &g,
                 std::string &err)
{
    const int regs_per_thread = TB == 2 ? 80 : 128;
    const char *env_odd = getenv("QMCB_ODD_ROWS");
    const int nbp = (env_odd && atoi(env_odd)) ? M.nb : M.nb + (M.nb & 1);
    double best = -1.0;
    const int kfull = M.kmax + 1;
    const char *env_kc = getenv("QMCB_KC");
    const char *env_nt = getenv("QMCB_NT");
    const char *env_il = getenv("QMCB_INTERLEAVE");
    const bool want_il = !env_il || atoi(env_il) != 0;
    for (int nt = 64; nt <= 256; nt += 32) {
        if (env_nt && atoi(env_nt) != nt) continue;
        int G = nt / M.nb;
        if (G < 1) continue;
        double eff = (double) (G * M.nb) / nt;
        int by_regs = 65536 / (regs_per_thread * nt);
        int by_thr = 2048 / nt;
        std::vector<int> kcs;
        if (env_kc)
            kcs.push_back(std::max(1, std::min(kfull, atoi(env_kc))));
        else
            for (int kc = kfull; kc >= 1; kc = (kc > 4 ? (kc + 1) / 2 : kc - 1))
                kcs.push_back(kc);
        for (int kc : kcs) {
            const bool il = want_il && G > 1 && (G & 1);
            int bytes = group_smem_doubles(G, nbp, M.nb, kc, il) * 8;
            if (bytes > max_smem) continue;
            int by_smem = (228 * 1024) / (bytes + 1024);
            int ctas = std::min(by_smem, std::min(by_regs, by_thr));
            if (ctas < 1) continue;
            double warps = ctas * nt / 32.0;
            // every halving of kc costs two more CTA barriers per walker
            double sync_pen = 1.0 - 0.004 * ((kfull + kc - 1) / kc - 1);
            double score = eff * std::min(1.0, warps / 16.0) * sync_pen;
            if (score > best + 1e-9) {
                best = score;
                g.nthreads = nt; g.G = G; g.nbp = nbp; g.kc = kc;
                g.interleave = il ? 1 : 0;
                g.tab_stride = group_tab_stride(nbp, M.nb, G, il);
                g.q_stride = group_q_stride(nbp, M.nb, kc, G, il);
                g.smem_bytes = bytes;
            }
        }
    }
    if (best < 0) {
        err = "boson_number too large for the shared-memory pair tables";
        return false;
    }
    return true;
}

// Kernel launch, optionally as a programmatic dependent of the kernel before
// it in the stream (cudaLaunchAttributeProgrammaticStreamSerialization).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block,
                       size_t smem, cudaStream_t stream, bool pdl,
                       Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename K>
int set_smem(qmcb_handle *h, K kernel)
{
    CUDA_TRY(h, cudaFuncSetAttribute(
                    kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                    h->geom.smem_bytes));
    return QMCB_OK;
}

int ensure_scratch(qmcb_handle *h, size_t bytes)
{
    if (bytes <= h->scratch_bytes) return QMCB_OK;
    if (h->d_scratch) CUDA_TRY(h, cudaFree(h->d_scratch));
    h->d_scratch = nullptr;
    h->scratch_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->d_scratch, bytes));
    h->scratch_bytes = bytes;
    return QMCB_OK;
}

int launch_model_eval(qmcb_handle *h, const EvalArgs &a, bool want_ln,
                      bool want_ef)
{
    if (a.nconf <= 0) return QMCB_OK;
    long long ctas = (a.nconf + h->geom.G - 1) / h->geom.G;
    long long cap = (long long) h->sm_count * 32;
    int grid = (int) std::min(ctas, cap);
    dim3 blk(h->geom.nthreads);
    size_t sm = h->geom.smem_bytes;
    if (want_ln && want_ef)
        model_eval_kernel<true, true><<<grid, blk, sm, h->stream>>>(
            h->M, h->geom, a);
    else if (want_ln)
        model_eval_kernel<true, false><<<grid, blk, sm, h->stream>>>(
            h->M, h->geom, a);
    else
        model_eval_kernel<false, true><<<grid, blk, sm, h->stream>>>(
            h->M, h->geom, a);
    CUDA_TRY(h, cudaGetLastError());
    return QMCB_OK;
}

void drop_block_graph(qmcb_handle *h)
{
    if (h->block_graph) cudaGraphExecDestroy(h->block_graph);
    h->block_graph = nullptr;
    h->block_graph_nts = 0;
}

void free_dmc(qmcb_handle *h)
{
    drop_block_graph(h);
    DmcBufs &B = h->B;
    for (int i = 0; i < 2; ++i) {
        cudaFree(B.confs[i]); cudaFree(B.energy[i]); cudaFree(B.weight[i]);
    }
    cudaFree(B.slot_energy); cudaFree(B.ref); cudaFree(B.cnt);
    cudaFree(B.blocksum); cudaFree(B.epart);
    cudaFree(B.ctl);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->ssf_aux[i]); cudaFree(h->den_hist[i]);
        h->ssf_aux[i] = nullptr; h->den_hist[i] = nullptr;
    }
    cudaFree(h->ssf_iter); cudaFree(h->den_total); cudaFree(h->den_iter);
    cudaFree(h->est_partial); cudaFree(h->den_hi);
    cudaFree(h->den_lists); cudaFree(h->den_wrec);
    h->ssf_iter = h->den_total = h->den_iter = h->est_partial = nullptr;
    h->den_hi = nullptr;
    h->den_lists = nullptr; h->den_wrec = nullptr;
    h->den_lists_steps = 0;
    h->est_log_cap = 0;
    cudaFree(h->L.energy); cudaFree(h->L.weight); cudaFree(h->L.ref_energy);
    cudaFree(h->L.accum_energy); cudaFree(h->L.num_walkers);
    cudaFree(h->X.aglob); cudaFree(h->X.slab); cudaFree(h->X.gathered);
    cudaFree(h->X.vred); cudaFree(h->X.offs);
    h->X = DmcMulti{};
    h->multi_ready = false;
    B = DmcBufs{};
    h->L = DmcLog{};
    h->log_cap = 0;
    h->dmc_ready = false;
}

void free_vmc(qmcb_handle *h)
{
    cudaFree(h->V.confs); cudaFree(h->V.lnpsi); cudaFree(h->V.eprev);
    cudaFree(h->V.ssfprev);
    cudaFree(h->vmc_sum_e); cudaFree(h->vmc_sum_ssf); cudaFree(h->vmc_acc);
    h->V = VmcState{};
    h->vmc_sum_e = h->vmc_sum_ssf = h->vmc_acc = nullptr;
    h->vmc_ready = false;
    h->vmc_chains = 0;
}

int alloc_dmc(qmcb_handle *h, const qmcb_dmc_params *p,
              long long slot_offset)
{
    const DmcConsts old_c = h->C;
    if (p->max_num_walkers < 1 || p->target_num_walkers < 1)
        FAIL(h, QMCB_ERR_INVALID, "max/target_num_walkers must be >= 1");
    long long cap = p->local_capacity > 0 ? p->local_capacity
                                          : p->max_num_walkers;
    if (cap > (1ll << 30)) FAIL(h, QMCB_ERR_INVALID, "capacity too large");
    if (!(p->time_step > 0)) FAIL(h, QMCB_ERR_INVALID, "time_step must be > 0");
    if (!(p->upper_bound > p->lower_bound))
        FAIL(h, QMCB_ERR_INVALID, "upper_bound must exceed lower_bound");
    DmcBufs &B = h->B;
    const int N = h->M.nop;
    size_t cb = (size_t) cap * 2 * N * sizeof(double);
    // a restart with the same capacity keeps the device allocation
    if (p->ssf_num_modes < 0 || p->density_num_bins < 0
        || p->ssf_num_modes > 65536 || p->density_num_bins > (1 << 24))
        FAIL(h, QMCB_ERR_INVALID, "bad estimator sizes");
    const bool reuse = B.confs[0] != nullptr && B.cap == (int) cap
                       && h->dp.ssf_num_modes == p->ssf_num_modes
                       && h->dp.density_num_bins == p->density_num_bins
                       && h->dp.density_as_pure == p->density_as_pure;
    if (!reuse) {
        free_dmc(h);
        B.cap = (int) cap;
        B.nblk = (int) ((cap + BR_TILE - 1) / BR_TILE);
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(h, cudaMalloc(&B.confs[i], cb));
            CUDA_TRY(h, cudaMalloc(&B.energy[i], cap * sizeof(double)));
            CUDA_TRY(h, cudaMalloc(&B.weight[i], cap * sizeof(double)));
        }
        CUDA_TRY(h, cudaMalloc(&B.slot_energy, cap * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&B.ref, cap * sizeof(int)));
        CUDA_TRY(h, cudaMalloc(&B.cnt, cap * sizeof(int)));
        CUDA_TRY(h, cudaMalloc(&B.blocksum, B.nblk * sizeof(long long)));
        CUDA_TRY(h, cudaMalloc(&B.epart, B.nblk * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&B.ctl, sizeof(DmcCtl)));
        for (int i = 0; i < 2; ++i)
            CUDA_TRY(h, cudaMemsetAsync(B.confs[i], 0, cb, h->stream));
        const size_t M3 = (size_t) p->ssf_num_modes * 3;
        const size_t NB = (size_t) p->density_num_bins;
        if (M3)
            for (int i = 0; i < 2; ++i)
                CUDA_TRY(h, cudaMalloc(&h->ssf_aux[i],
                                       cap * M3 * sizeof(double)));
        // 16-bit bin lists instead of per-slot histograms
        // (QMCB_DENSITY_HIST keeps the histogram path, for A/B tests)
        h->den_lists_mode = NB && NB <= 65536
                            && !getenv("QMCB_DENSITY_HIST");
        if (NB) {
            const int nh = h->den_lists_mode ? 0
                           : (p->density_as_pure ? 1 : 2);
            for (int i = 0; i < nh; ++i)
                CUDA_TRY(h, cudaMalloc(&h->den_hist[i],
                                       cap * NB * sizeof(double)));
            CUDA_TRY(h, cudaMalloc(&h->den_total, 3 * NB * sizeof(double)));
            CUDA_TRY(h, cudaMalloc(&h->den_hi, sizeof(int)));
        }
        if (M3 || NB)
            CUDA_TRY(h, cudaMalloc(&h->est_partial,
                                   CS_BLOCKS * std::max(M3, NB)
                                       * sizeof(double)));
    }
    // the bin lists of the pure density estimator are sized by the
    // forward-walking window
    if (h->dp.density_pfw_nts != p->density_pfw_nts) h->est_log_cap = 0;
    h->dmc_ready = false;
    h->dp = *p;
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(h, cudaMemsetAsync(B.energy[i], 0, cap * sizeof(double),
                                    h->stream));
        CUDA_TRY(h, cudaMemsetAsync(B.weight[i], 0, cap * sizeof(double),
                                    h->stream));
    }
    CUDA_TRY(h, cudaMemsetAsync(B.slot_energy, 0, cap * sizeof(double),
                                h->stream));
    CUDA_TRY(h, cudaMemsetAsync(B.ref, 0, cap * sizeof(int), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(B.ctl, 0, sizeof(DmcCtl), h->stream));

    DmcConsts &C = h->C;
    C.dt = p->time_step;
    C.sigma = std::sqrt(2 * p->time_step);   // mrbp_qmc/dmc.py:178
    C.z_min = p->lower_bound;
    C.size = p->upper_bound - p->lower_bound;
    C.nwc_over_dt = p->nwc_factor / p->time_step;
    C.target = (double) p->target_num_walkers;
    C.seed = p->rng_seed;
    C.slot_offset = slot_offset;
    C.energy_mode = p->energy_mode;
    C.defer_weight = (h->comm && p->energy_mode == 0) ? 1 : 0;
    h->weights_pending = false;
    h->multi_ready = false;     // the global stale-slot array is rebuilt from
                                // the state about to be loaded
    // the nodes of the block graph hold the constants by value (a
    // reallocation has already dropped it in free_dmc)
    if (old_c.dt != C.dt || old_c.sigma != C.sigma || old_c.z_min != C.z_min
        || old_c.size != C.size || old_c.nwc_over_dt != C.nwc_over_dt
        || old_c.target != C.target || old_c.seed != C.seed
        || old_c.slot_offset != C.slot_offset
        || old_c.energy_mode != C.energy_mode
        || old_c.defer_weight != C.defer_weight)
        drop_block_graph(h);
    return QMCB_OK;
}

int ensure_est_log(qmcb_handle *h, long long nts)
{
    if (nts <= h->est_log_cap) return QMCB_OK;
    cudaFree(h->ssf_iter); cudaFree(h->den_iter);
    h->ssf_iter = h->den_iter = nullptr;
    h->est_log_cap = 0;
    const size_t M3 = (size_t) h->dp.ssf_num_modes * 3;
    const size_t NB = (size_t) h->dp.density_num_bins;
    if (M3) CUDA_TRY(h, cudaMalloc(&h->ssf_iter, nts * M3 * sizeof(double)));
    if (NB) CUDA_TRY(h, cudaMalloc(&h->den_iter, nts * NB * sizeof(double)));
    if (h->den_lists_mode) {
        cudaFree(h->den_lists); cudaFree(h->den_wrec);
        h->den_lists = nullptr; h->den_wrec = nullptr;
        // pure: the steps of the forward-walking window; mixed: all steps
        const long long steps = h->dp.density_as_pure
            ? std::min<long long>(
                  nts, std::max<long long>(1, h->dp.density_pfw_nts))
            : nts;
        CUDA_TRY(h, cudaMalloc(&h->den_lists,
                               (size_t) steps * h->B.cap * h->M.nop
                                   * sizeof(unsigned short)));
        CUDA_TRY(h, cudaMalloc(&h->den_wrec, steps * sizeof(int)));
        h->den_lists_steps = steps;
    }
    h->est_log_cap = nts;
    return QMCB_OK;
}

// rho_k of `rows` configurations: staged-seed kernel when a walker's chunks
// fit one CTA, the exact-seed kernel otherwise (M > 2048).
int launch_ssf_eval(qmcb_handle *h, const SsfArgs &a, long long rows)
{
    const int N = a.N, M = a.M;
    const int nchunk = (M + SSF_CHUNK - 1) / SSF_CHUNK;
    static const bool force_exact = getenv("QMCB_SSF_EXACT_SEEDS") != nullptr;
    if (nchunk <= SSF_THREADS && !force_exact) {
        SsfStage st{};
        st.nchunk = nchunk;
        st.G = std::max(1, SSF_THREADS / nchunk);
        // ~40 KB of seeds per CTA (5 CTAs per SM next to 96 registers)
        const size_t budget = 40 * 1024;
        st.P = (int) (budget / ((size_t) st.G * (nchunk + 1) * 16));
        st.P = std::max(1, std::min(st.P, std::min(N, 32)));
        if (const char *ep = getenv("QMCB_SSF_P"))
            st.P = std::max(1, std::min(atoi(ep), N));
        while (st.G > 1 && ssf_stage_smem(st.G, nchunk, st.P) > 96 * 1024)
            --st.G;
        size_t smem = ssf_stage_smem(st.G, nchunk, st.P);
        if (smem > 48 * 1024)
            CUDA_TRY(h, cudaFuncSetAttribute(
                            ssf_eval_staged_kernel,
                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int) smem));
        int grid = (int) std::min<long long>((rows + st.G - 1) / st.G,
                                             (long long) h->sm_count * 16);
        ssf_eval_staged_kernel<<<std::max(grid, 1), SSF_THREADS, smem,
                                 h->stream>>>(a, st);
    } else {
        int G = std::max(1, SSF_THREADS / nchunk);
        G = std::min(G, std::max(1, (40 * 1024) / (24 * N)));
        size_t smem = (size_t) 3 * G * N * sizeof(double);
        int grid = (int) std::min<long long>((rows + G - 1) / G,
                                             (long long) h->sm_count * 16);
        ssf_eval_kernel<<<std::max(grid, 1), SSF_THREADS, smem, h->stream>>>(
            a, G, nchunk);
    }
    CUDA_TRY(h, cudaGetLastError());
    return QMCB_OK;
}

// One step of the S(k) estimator on the "actual" population
// (qmc_base/jastrow/dmc.py:363-573): step_idx counts from the block start.
int launch_ssf_step(qmcb_handle *h, long long step_idx)
{
    DmcBufs &B = h->B;
    const int M = h->dp.ssf_num_modes, N = h->M.nop;
    const bool pure = h->dp.ssf_as_pure != 0;
    const long long pfw = h->dp.ssf_pfw_nts;
    const int cur = (int) (step_idx & 1);
    int *W_dev = (int *) ((char *) B.ctl + offsetof(DmcCtl, W));
    double *out = h->ssf_aux[cur];
    const double *prev = h->ssf_aux[cur ^ 1];
    if (pure && step_idx >= pfw) {
        rows_gather_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(
            prev, B.ref, W_dev, 3 * M, out);
    } else {
        SsfArgs a{};
        // parent buffer of the step in flight: parity of ctl->step, which
        // the host mirrors in step_host
        a.confs = B.confs[(int) ((h->step_host + step_idx) & 1)];
        a.ref = B.ref; a.W_dev = W_dev; a.N = N; a.M = M;
        a.two_over_L = 2.0 / h->M.L;
        a.out = out; a.prev = prev; a.accumulate = pure ? 1 : 0;
        int rc = launch_ssf_eval(h, a, B.cap);
        if (rc) return rc;
    }
    RowRange rr{};
    rr.hi_dev = W_dev; rr.lo_host = 0;
    colsum_partial_kernel<<<CS_BLOCKS, 256, 0, h->stream>>>(
        out, rr, 3 * M, h->est_partial);
    double div = 1.0;
    if (pure) div = (double) std::min(step_idx + 1, pfw);
    colsum_final_kernel<<<(3 * M + 127) / 128, 128, 0, h->stream>>>(
        h->est_partial, CS_BLOCKS, 3 * M, nullptr, 1.0, 1.0 / div,
        h->ssf_iter + step_idx * 3 * M);
    CUDA_TRY(h, cudaGetLastError());
    return QMCB_OK;
}

// One step of the density estimator (mrbp_qmc/dmc.py:472-547).  Pure mode
// ignores the genealogy (quirk Q2), so the reference's ping-pong copy of all
// slots is the same as one in-place per-slot histogram; mixed mode keeps the
// two parity buffers the reference accumulates into (quirk Q3).
int launch_density_step(qmcb_handle *h, long long step_idx)
{
    DmcBufs &B = h->B;
    const int NB = h->dp.density_num_bins, N = h->M.nop;
    const bool pure = h->dp.density_as_pure != 0;
    const long long pfw = h->dp.density_pfw_nts;
    const int pbuf = pure ? 0 : (int) (step_idx & 1);
    int *W_dev = (int *) ((char *) B.ctl + offsetof(DmcCtl, W));
    if (h->den_lists_mode) {
        // mixed mode: one running total per step parity, every step recorded
        double *total = h->den_total + (size_t) pbuf * NB;
        double *corr = h->den_total + 2 * (size_t) NB;
        const long long stride = (long long) B.cap * N;
        const bool record = pure ? step_idx < std::min(pfw, h->den_lists_steps)
                                 : step_idx < h->den_lists_steps;
        if (record) {
            const double *confs =
                B.confs[(int) ((h->step_host + step_idx) & 1)];
            const size_t sm_bytes = (size_t) NB * sizeof(unsigned int);
            const int use_smem = sm_bytes <= 160 * 1024;
            if (use_smem && sm_bytes > 48 * 1024)
                CUDA_TRY(h, cudaFuncSetAttribute(
                                density_list_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int) sm_bytes));
            density_list_kernel<<<h->sm_count * (use_smem ? 2 : 8), 256,
                                  use_smem ? sm_bytes : 0, h->stream>>>(
                confs, B.ref, W_dev, N, NB, h->M.L / NB, B.cap,
                h->den_lists + step_idx * stride, h->den_wrec + step_idx,
                total, use_smem);
        }
        // recorded steps whose counts in the slots beyond W_t must go
        long long nrec;
        int first = 0, every = 1;
        if (pure) {
            nrec = std::min<long long>(std::min(step_idx + 1, pfw),
                                       h->den_lists_steps);
        } else {
            first = pbuf; every = 2;
            nrec = std::min(step_idx, h->den_lists_steps - 1) / 2 + 1;
            if (first > step_idx) nrec = 0;
        }
        if (nrec > 0)
            density_corr_kernel<<<dim3((unsigned) nrec, 16), 256, 0,
                                  h->stream>>>(h->den_lists, h->den_wrec,
                                               W_dev, N, stride, first, every,
                                               corr);
        const double div = pure ? (double) std::max<long long>(
                                      1, std::min(step_idx + 1, pfw))
                                : 1.0;
        density_out_kernel<<<(NB + 127) / 128, 128, 0, h->stream>>>(
            total, corr, NB, 1.0 / div, h->den_iter + step_idx * NB);
        CUDA_TRY(h, cudaGetLastError());
        return QMCB_OK;
    }
    double *hist = h->den_hist[pbuf];
    double *total = h->den_total + (size_t) pbuf * NB;
    if (!pure || step_idx < pfw) {
        const double *confs = B.confs[(int) ((h->step_host + step_idx) & 1)];
        const size_t sm_bytes = (size_t) NB * sizeof(unsigned int);
        const int use_smem = sm_bytes <= 160 * 1024;
        if (use_smem && sm_bytes > 48 * 1024)
            CUDA_TRY(h, cudaFuncSetAttribute(
                            density_hist_kernel,
                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int) sm_bytes));
        density_hist_kernel<<<h->sm_count * (use_smem ? 2 : 8), 256,
                              use_smem ? sm_bytes : 0, h->stream>>>(
            confs, B.ref, W_dev, N, NB, h->M.L / NB, hist, total, h->den_hi,
            use_smem);
    }
    // sum over live slots = running total - rows of slots that died
    RowRange rr{};
    rr.lo_dev = W_dev; rr.hi_dev = h->den_hi;
    colsum_partial_kernel<<<CS_BLOCKS, 256, 0, h->stream>>>(
        hist, rr, NB, h->est_partial);
    double div = 1.0;
    if (pure) div = (double) std::min(step_idx + 1, pfw);
    colsum_final_kernel<<<(NB + 127) / 128, 128, 0, h->stream>>>(
        h->est_partial, CS_BLOCKS, NB, total, -1.0, 1.0 / div,
        h->den_iter + step_idx * NB);
    CUDA_TRY(h, cudaGetLastError());
    return QMCB_OK;
}

int ensure_log(qmcb_handle *h, long long nts)
{
    if (nts <= h->log_cap) return QMCB_OK;
    drop_block_graph(h);        // its nodes hold the old log pointers
    DmcLog &L = h->L;
    cudaFree(L.energy); cudaFree(L.weight); cudaFree(L.ref_energy);
    cudaFree(L.accum_energy); cudaFree(L.num_walkers);
    L = DmcLog{};
    h->log_cap = 0;
    CUDA_TRY(h, cudaMalloc(&L.energy, nts * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&L.weight, nts * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&L.ref_energy, nts * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&L.accum_energy, nts * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&L.num_walkers, nts * sizeof(unsigned long long)));
    h->log_cap = nts;
    return QMCB_OK;
}

int read_ctl(qmcb_handle *h, DmcCtl &ctl)
{
    CUDA_TRY(h, cudaMemcpyAsync(&ctl, h->B.ctl, sizeof ctl,
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

void fill_scalars(const qmcb_handle *h, const DmcCtl &ctl,
                  qmcb_state_scalars *s)
{
    if (!s) return;
    s->energy = ctl.last_energy;
    s->weight = ctl.last_weight;
    s->ref_energy = ctl.eref[ctl.step & 1];
    s->accum_energy = ctl.last_accum;
    s->total_energy = ctl.tot_e;
    s->total_weight = ctl.tot_w;
    s->num_walkers = ctl.step == 0 ? ctl.W_prev : ctl.W;
    s->max_num_walkers = h->B.cap;
    s->step = ctl.step;
    s->capacity_hits = ctl.capacity_hits;
}

// Sharded runs: form the branching weights the last step left pending and
// bring the global stale-energy array up to date (what the next branching
// would otherwise do on the fly).  Needed before anything reads or moves the
// weights: the end of a block, a rebalance.
int materialize_weights(qmcb_handle *h)
{
    if (!h->comm || !h->weights_pending) return QMCB_OK;
    const int grid = h->sm_count * 4;
    multi_weight_kernel<<<grid, 256, 0, h->stream>>>(h->B, h->C, h->X);
    CUDA_TRY(h, cudaEventRecord(h->ev_weighted, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->pc_stream, h->ev_weighted, 0));
    multi_apply_kernel<<<grid, 256, 0, h->pc_stream>>>(h->X);
    CUDA_TRY(h, cudaEventRecord(h->ev_controlled, h->pc_stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_controlled, 0));
    CUDA_TRY(h, cudaGetLastError());
    h->weights_pending = false;
    return QMCB_OK;
}

// Sharded runs: allocate the buffers of the per-step collectives and
// (re)build the replicated global stale-slot array from the local arrays of
// all ranks.  Import convention (qmcb_dmc_init / qmcb_dmc_set_state): rank q
// holds, at local index i < n_q, A at the global position o_q + i of its
// i-th walker; the entries of the LAST rank beyond its population are the
// tail of the global array (positions beyond the live ensemble).
int multi_setup(qmcb_handle *h)
{
    if (!h->comm || h->multi_ready) return QMCB_OK;
    NcclApi *api = nccl_api();
    DmcBufs &B = h->B;
    DmcMulti &X = h->X;
    const int R = h->world, me = h->rank;
    if (!X.aglob || X.cap != B.cap || X.R != R) {
        cudaFree(X.aglob); cudaFree(X.slab); cudaFree(X.gathered);
        cudaFree(X.vred); cudaFree(X.offs);
        X = DmcMulti{};
        const size_t rc = (size_t) R * B.cap;
        CUDA_TRY(h, cudaMalloc(&X.aglob, rc * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&X.gathered, rc * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&X.slab, (size_t) B.cap * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&X.vred, (R + 2) * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&X.offs, (R + 1) * sizeof(long long)));
        X.R = R; X.rank = me; X.cap = B.cap;
    }
    DmcCtl ctl;
    int rc = read_ctl(h, ctl);
    if (rc) return rc;
    long long mine = ctl.W_prev;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_counts + me, &mine, sizeof mine,
                                cudaMemcpyHostToDevice, h->stream));
    NCCL_TRY(h, api->AllGather(h->d_counts + me, h->d_counts, 1, ncclInt64,
                               h->comm, h->stream));
    std::vector<long long> cnt(R), offs(R + 1, 0);
    CUDA_TRY(h, cudaMemcpyAsync(cnt.data(), h->d_counts,
                                R * sizeof(long long), cudaMemcpyDeviceToHost,
                                h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < R; ++q) offs[q + 1] = offs[q] + cnt[q];
    CUDA_TRY(h, cudaMemcpyAsync(X.offs, offs.data(),
                                (R + 1) * sizeof(long long),
                                cudaMemcpyHostToDevice, h->stream));
    long long pb[2] = {offs[me], offs[me]};
    CUDA_TRY(h, cudaMemcpyAsync(
                    (char *) B.ctl + offsetof(DmcCtl, pos_base), pb, sizeof pb,
                    cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(X.aglob, 0,
                                (size_t) R * B.cap * sizeof(double),
                                h->stream));
    NCCL_TRY(h, api->AllGather(B.slot_energy, X.gathered, B.cap, ncclDouble,
                               h->comm, h->stream));
    const int grid = h->sm_count * 4;
    multi_apply_kernel<<<grid, 256, 0, h->stream>>>(X);
    multi_tail_kernel<<<grid, 256, 0, h->stream>>>(X, cnt[R - 1]);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->multi_ready = true;
    return QMCB_OK;
}

}  // namespace

extern "C" {

const char *qmcb_version(void) { return "qmcb200 0.1.0 (sm_100a)"; }

const char *qmcb_last_error(const qmcb_handle *h)
{
    return h ? h->err.c_str() : g_create_error.c_str();
}

int qmcb_create(const qmcb_model_params *params, int device,
                qmcb_handle **out)
{
    if (!params || !out) {
        g_create_error = "null argument";
        return QMCB_ERR_INVALID;
    }
    *out = nullptr;
    qmcb_handle *h = new qmcb_handle();
    auto fail = [&](int code) {
        g_create_error = h->err;
        if (h->stream) cudaStreamDestroy(h->stream);
        delete h;
        return code;
    };
    h->device = device;
    h->params = *params;
    if (!build_model(*params, h->M, h->err)) return fail(QMCB_ERR_INVALID);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        h->err = std::string("no CUDA device available: ")
                 + cudaGetErrorString(e)
                 + " (this engine has no CPU fallback)";
        return fail(QMCB_ERR_CUDA);
    }
    if (device < 0 || device >= ndev) {
        h->err = "device index out of range";
        return fail(QMCB_ERR_INVALID);
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        return fail(QMCB_ERR_CUDA);
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        return fail(QMCB_ERR_CUDA);
    }
    h->sm_count = prop.multiProcessorCount;
    h->max_smem = (int) prop.sharedMemPerBlockOptin;
    if (!choose_geom(h->M, h->max_smem, h->geom, h->err))
        return fail(QMCB_ERR_INVALID);
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking))
        != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        return fail(QMCB_ERR_CUDA);
    }
    int rc;
    if ((rc = set_smem(h, model_eval_kernel<true, true>)) != QMCB_OK
        || (rc = set_smem(h, model_eval_kernel<true, false>)) != QMCB_OK
        || (rc = set_smem(h, model_eval_kernel<false, true>)) != QMCB_OK
        || (rc = set_smem(h, dmc_step_kernel<false>)) != QMCB_OK
        || (rc = set_smem(h, dmc_step_kernel<true>)) != QMCB_OK
        || (rc = build_trig_tables(h)) != QMCB_OK)
        return fail(rc);
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    *out = h;
    return QMCB_OK;
}

void qmcb_destroy(qmcb_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_dmc(h);
    free_vmc(h);
    cudaFree(h->d_scratch);
    cudaFree(h->d_trig_z); cudaFree(h->d_trig_c);
    cudaFree(h->d_counts);
    cudaFree(h->rb.sum); cudaFree(h->rb.sqr); cudaFree(h->rb.nblk);
    cudaFree(h->cs_confs); cudaFree(h->cs_ln0); cudaFree(h->cs_work);
    if (h->pc_stream) {
        cudaStreamSynchronize(h->pc_stream);
        cudaStreamDestroy(h->pc_stream);
    }
    if (h->ev_branched) cudaEventDestroy(h->ev_branched);
    if (h->ev_controlled) cudaEventDestroy(h->ev_controlled);
    if (h->ev_weighted) cudaEventDestroy(h->ev_weighted);
    if (h->comm && nccl_api()) nccl_api()->CommDestroy(h->comm);
    for (auto ev : h->step_ev) cudaEventDestroy(ev);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
}

int qmcb_model_eval_device(qmcb_handle *h, const double *d_confs,
                           int64_t nconf, double *d_lnpsi, double *d_energy,
                           double *d_drift)
{
    if (!h) return QMCB_ERR_INVALID;
    if (nconf < 0) FAIL(h, QMCB_ERR_INVALID, "nconf < 0");
    CUDA_TRY(h, cudaSetDevice(h->device));
    EvalArgs a{};
    a.confs = d_confs; a.nconf = nconf;
    a.lnpsi = d_lnpsi; a.energy = d_energy; a.drift = d_drift;
    bool ln = d_lnpsi != nullptr, ef = d_energy || d_drift;
    if (!ln && !ef) return QMCB_OK;
    return launch_model_eval(h, a, ln, ef);
}

int qmcb_model_eval(qmcb_handle *h, const double *confs, int64_t nconf,
                    double *lnpsi, double *energy, double *drift)
{
    if (!h) return QMCB_ERR_INVALID;
    if (nconf < 0 || (nconf > 0 && !confs))
        FAIL(h, QMCB_ERR_INVALID, "bad confs / nconf");
    if (nconf == 0) return QMCB_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = h->M.nop;
    size_t nc = (size_t) nconf * 2 * N, nd = (size_t) nconf * N;
    size_t total = (nc + nd + 2 * (size_t) nconf) * sizeof(double);
    int rc = ensure_scratch(h, total);
    if (rc) return rc;
    double *d_confs = h->d_scratch, *d_drift = d_confs + nc;
    double *d_ln = d_drift + nd, *d_e = d_ln + nconf;
    CUDA_TRY(h, cudaMemcpyAsync(d_confs, confs, nc * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    rc = qmcb_model_eval_device(h, d_confs, nconf, lnpsi ? d_ln : nullptr,
                                energy ? d_e : nullptr,
                                drift ? d_drift : nullptr);
    if (rc) return rc;
    if (lnpsi)
        CUDA_TRY(h, cudaMemcpyAsync(lnpsi, d_ln, nconf * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (energy)
        CUDA_TRY(h, cudaMemcpyAsync(energy, d_e, nconf * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (drift)
        CUDA_TRY(h, cudaMemcpyAsync(drift, d_drift, nd * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

int qmcb_fourier_density(qmcb_handle *h, const double *confs, int64_t nconf,
                         int32_t num_modes, double *out)
{
    if (!h) return QMCB_ERR_INVALID;
    if (nconf < 0 || num_modes < 1 || (nconf > 0 && (!confs || !out)))
        FAIL(h, QMCB_ERR_INVALID, "bad arguments");
    if (nconf == 0) return QMCB_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = h->M.nop, M = num_modes;
    size_t nc = (size_t) nconf * 2 * N, no = (size_t) nconf * M * 3;
    int rc = ensure_scratch(h, (nc + no) * sizeof(double));
    if (rc) return rc;
    double *d_confs = h->d_scratch, *d_out = d_confs + nc;
    CUDA_TRY(h, cudaMemcpyAsync(d_confs, confs, nc * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    SsfArgs a{};
    a.confs = d_confs; a.W_host = nconf; a.N = N; a.M = M;
    a.two_over_L = 2.0 / h->M.L;
    a.out = d_out;
    rc = launch_ssf_eval(h, a, nconf);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(out, d_out, no * sizeof(double),
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

int qmcb_one_body_density_device(qmcb_handle *h, const double *d_confs,
                                 int64_t nconf, const double *d_offsets,
                                 int32_t num_offsets, double *d_out)
{
    if (!h) return QMCB_ERR_INVALID;
    if (nconf < 0 || num_offsets < 0
        || (nconf > 0 && num_offsets > 0
            && (!d_confs || !d_offsets || !d_out)))
        FAIL(h, QMCB_ERR_INVALID, "bad arguments");
    if (nconf == 0 || num_offsets == 0) return QMCB_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = h->M.nop, S = num_offsets;
    // items (configuration, offset) are dealt to CTAs in order; a CTA needs
    // the tables of every configuration its items touch
    int nt = 128, per_conf = 0;
    size_t smem = 0;
    for (; nt >= 32; nt /= 2) {
        int gmax = (nt - 1) / S + 2;
        if ((long long) gmax > nconf) gmax = (int) nconf;
        smem = (size_t) gmax * obd_slot_doubles(N) * sizeof(double);
        if (smem <= (size_t) h->max_smem) break;
    }
    if (nt < 32) {
        // large N: the tables of a single configuration per CTA
        nt = 128;
        per_conf = 1;
        smem = obd_slot_doubles(N) * sizeof(double);
        if (smem > (size_t) h->max_smem)
            FAIL(h, QMCB_ERR_INVALID,
                 "boson_number too large for the one-body density tables");
    }
    if (smem > 48 * 1024)
        CUDA_TRY(h, cudaFuncSetAttribute(
                        obd_kernel,
                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                        (int) smem));
    ObdArgs a{};
    a.confs = d_confs; a.nconf = nconf; a.offsets = d_offsets; a.S = S;
    a.out = d_out;
    a.per_conf = per_conf;
    long long items = (long long) nconf * S;
    long long grid = per_conf ? (long long) nconf * ((S + nt - 1) / nt)
                              : (items + nt - 1) / nt;
    if (grid > 0x7fffffffll)
        FAIL(h, QMCB_ERR_INVALID, "too many (configuration, offset) items");
    obd_kernel<<<(unsigned) grid, nt, smem, h->stream>>>(h->M, a);
    CUDA_TRY(h, cudaGetLastError());
    return QMCB_OK;
}

int qmcb_one_body_density(qmcb_handle *h, const double *confs, int64_t nconf,
                          const double *offsets, int32_t num_offsets,
                          double *out)
{
    if (!h) return QMCB_ERR_INVALID;
    if (nconf < 0 || num_offsets < 0
        || (nconf > 0 && num_offsets > 0 && (!confs || !offsets || !out)))
        FAIL(h, QMCB_ERR_INVALID, "bad arguments");
    if (nconf == 0 || num_offsets == 0) return QMCB_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = h->M.nop;
    size_t nc = (size_t) nconf * 2 * N, no = (size_t) nconf * num_offsets;
    int rc = ensure_scratch(h, (nc + no + num_offsets) * sizeof(double));
    if (rc) return rc;
    double *d_confs = h->d_scratch, *d_out = d_confs + nc, *d_off = d_out + no;
    CUDA_TRY(h, cudaMemcpyAsync(d_confs, confs, nc * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_off, offsets, num_offsets * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    rc = qmcb_one_body_density_device(h, d_confs, nconf, d_off, num_offsets,
                                      d_out);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(out, d_out, no * sizeof(double),
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

int qmcb_fourier_density_k(qmcb_handle *h, const double *confs,
                           int64_t nconf, const double *kz, int32_t nk,
                           double *out)
{
    if (!h) return QMCB_ERR_INVALID;
    if (nconf < 0 || nk < 0
        || (nconf > 0 && nk > 0 && (!confs || !kz || !out)))
        FAIL(h, QMCB_ERR_INVALID, "bad arguments");
    if (nconf == 0 || nk == 0) return QMCB_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = h->M.nop;
    size_t nc = (size_t) nconf * 2 * N, no = (size_t) nconf * nk * 2;
    int rc = ensure_scratch(h, (nc + no + nk) * sizeof(double));
    if (rc) return rc;
    double *d_confs = h->d_scratch, *d_out = d_confs + nc, *d_k = d_out + no;
    CUDA_TRY(h, cudaMemcpyAsync(d_confs, confs, nc * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_k, kz, nk * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    long long items = (long long) nconf * nk;
    int grid = (int) std::min<long long>((items + 127) / 128,
                                         (long long) h->sm_count * 32);
    fdk_general_kernel<<<grid, 128, 0, h->stream>>>(d_confs, nconf, N, d_k,
                                                    nk, d_out);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(out, d_out, no * sizeof(double),
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

int qmcb_set_model_params(qmcb_handle *h, const qmcb_model_params *params)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!params) FAIL(h, QMCB_ERR_INVALID, "null params");
    DevModel M{};
    std::string err;
    if (!build_model(*params, M, err)) FAIL(h, QMCB_ERR_INVALID, err);
    if (M.nop != h->M.nop)
        FAIL(h, QMCB_ERR_INVALID,
             "boson_number cannot change on a live handle (the launch "
             "geometry is fixed at qmcb_create)");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // kernels in flight hold their own copy of the constants (passed by
    // value), so no synchronisation is needed
    // the node tables are rebuilt in place: wait for their readers
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->params = *params;
    h->M = M;
    drop_block_graph(h);    // the step kernel's constants are node arguments
    return build_trig_tables(h);
}

int qmcb_cs_load(qmcb_handle *h, const double *confs, int64_t nconf,
                 const double *ini_lnpsi)
{
    if (!h) return QMCB_ERR_INVALID;
    if (nconf < 1 || !confs)
        FAIL(h, QMCB_ERR_INVALID, "bad confs / nconf");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = h->M.nop;
    cudaFree(h->cs_confs); cudaFree(h->cs_ln0); cudaFree(h->cs_work);
    h->cs_confs = h->cs_ln0 = h->cs_work = nullptr;
    h->cs_n = 0;
    size_t nc = (size_t) nconf * 2 * N * sizeof(double);
    CUDA_TRY(h, cudaMalloc(&h->cs_confs, nc));
    CUDA_TRY(h, cudaMalloc(&h->cs_ln0, nconf * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&h->cs_work, (2 * nconf + 4) * sizeof(double)));
    CUDA_TRY(h, cudaMemcpyAsync(h->cs_confs, confs, nc,
                                cudaMemcpyHostToDevice, h->stream));
    if (ini_lnpsi) {
        CUDA_TRY(h, cudaMemcpyAsync(h->cs_ln0, ini_lnpsi,
                                    nconf * sizeof(double),
                                    cudaMemcpyHostToDevice, h->stream));
    } else {
        // the set was drawn from the handle's current wave function
        EvalArgs a{};
        a.confs = h->cs_confs; a.nconf = nconf; a.lnpsi = h->cs_ln0;
        int rc = launch_model_eval(h, a, true, false);
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->cs_n = nconf;
    return QMCB_OK;
}

int qmcb_cs_variance(qmcb_handle *h, const qmcb_model_params *trial,
                     double *variance, double *ref_energy, double *lnpsi,
                     double *energy)
{
    if (!h) return QMCB_ERR_INVALID;
    if (h->cs_n < 1) FAIL(h, QMCB_ERR_STATE, "qmcb_cs_load not called");
    CUDA_TRY(h, cudaSetDevice(h->device));
    DevModel M = h->M;
    if (trial) {
        std::string err;
        if (!build_model(*trial, M, err)) FAIL(h, QMCB_ERR_INVALID, err);
        if (M.nop != h->M.nop)
            FAIL(h, QMCB_ERR_INVALID, "trial boson_number differs");
    }
    const long long n = h->cs_n;
    double *d_ln = h->cs_work, *d_e = d_ln + n, *d_res = d_e + n;
    {
        EvalArgs a{};
        a.confs = h->cs_confs; a.nconf = n; a.lnpsi = d_ln; a.energy = d_e;
        long long ctas = (n + h->geom.G - 1) / h->geom.G;
        int grid = (int) std::min(ctas, (long long) h->sm_count * 32);
        model_eval_kernel<true, true>
            <<<grid, h->geom.nthreads, h->geom.smem_bytes, h->stream>>>(
                M, h->geom, a);
    }
    cs_variance_kernel<<<1, CS_THREADS, 0, h->stream>>>(d_ln, h->cs_ln0, d_e,
                                                        n, d_res);
    CUDA_TRY(h, cudaGetLastError());
    double res[4];
    CUDA_TRY(h, cudaMemcpyAsync(res, d_res, sizeof res,
                                cudaMemcpyDeviceToHost, h->stream));
    if (lnpsi)
        CUDA_TRY(h, cudaMemcpyAsync(lnpsi, d_ln, n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (energy)
        CUDA_TRY(h, cudaMemcpyAsync(energy, d_e, n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (variance) *variance = res[0];
    if (ref_energy) *ref_energy = res[1];
    return QMCB_OK;
}

int qmcb_dmc_init(qmcb_handle *h, const qmcb_dmc_params *params,
                  const double *ini_confs, int64_t n, double ref_energy,
                  int64_t global_slot_offset)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!params || (n > 0 && !ini_confs) || n < 0)
        FAIL(h, QMCB_ERR_INVALID, "bad arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = alloc_dmc(h, params, global_slot_offset);
    if (rc) return rc;
    if (n > h->B.cap)
        FAIL(h, QMCB_ERR_INVALID, "more initial walkers than capacity");
    const int N = h->M.nop;
    DmcBufs &B = h->B;
    if (n > 0) {
        size_t nc = (size_t) n * 2 * N * sizeof(double);
        rc = ensure_scratch(h, nc);
        if (rc) return rc;
        CUDA_TRY(h, cudaMemcpyAsync(h->d_scratch, ini_confs, nc,
                                    cudaMemcpyHostToDevice, h->stream));
        EvalArgs a{};
        a.confs = h->d_scratch; a.nconf = n;
        a.energy = B.energy[0];
        a.state_confs = B.confs[0];
        a.state_weight = B.weight[0];
        a.slot_energy = B.slot_energy;
        rc = launch_model_eval(h, a, false, true);
        if (rc) return rc;
    }
    // initial scalars: mrbp_qmc/dmc.py:299-312
    std::vector<double> e(n);
    if (n > 0)
        CUDA_TRY(h, cudaMemcpyAsync(e.data(), B.energy[0], n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    double se = 0.0;
    for (int64_t i = 0; i < n; ++i) se += e[i];
    double n_glob = (double) n, se_glob = se;
    if (h->comm) {
        // the initial reference energy is the mean local energy of the
        // WHOLE ensemble (mrbp_qmc/dmc.py:299-312), identical on every rank
        double red[2] = {se, (double) n};
        double *d_red = (double *) ((char *) B.ctl + offsetof(DmcCtl, red));
        CUDA_TRY(h, cudaMemcpyAsync(d_red, red, sizeof red,
                                    cudaMemcpyHostToDevice, h->stream));
        NCCL_TRY(h, nccl_api()->AllReduce(d_red, d_red, 2, ncclDouble,
                                          ncclSum, h->comm, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(red, d_red, sizeof red,
                                    cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        se_glob = red[0];
        n_glob = red[1];
    }
    DmcCtl ctl{};
    ctl.W_prev = (int) n;
    ctl.W = (int) n;
    ctl.last_energy = se_glob;
    ctl.last_weight = n_glob;
    ctl.last_accum = n_glob > 0 ? se_glob / n_glob : 0.0;
    ctl.eref[0] = std::isnan(ref_energy) ? ctl.last_accum : ref_energy;
    ctl.eref[1] = ctl.eref[0];
    ctl.W_global = n_glob;
    ctl.pos_base[0] = ctl.pos_base[1] = global_slot_offset;
    CUDA_TRY(h, cudaMemcpyAsync(B.ctl, &ctl, sizeof ctl,
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->step_host = 0;
    h->n_ini = (int) n;
    h->dmc_ready = true;
    return QMCB_OK;
}

int qmcb_dmc_set_state(qmcb_handle *h, const qmcb_dmc_params *params,
                       const double *confs, const double *energy,
                       const double *weight, const double *slot_energy,
                       const qmcb_state_scalars *sc,
                       int64_t global_slot_offset)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!params || !sc || !confs || !energy || !weight)
        FAIL(h, QMCB_ERR_INVALID, "bad arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = alloc_dmc(h, params, global_slot_offset);
    if (rc) return rc;
    DmcBufs &B = h->B;
    int64_t n = sc->num_walkers;
    if (n < 0 || n > B.cap)
        FAIL(h, QMCB_ERR_INVALID, "num_walkers exceeds capacity");
    const int N = h->M.nop;
    int par = (int) (sc->step & 1);
    CUDA_TRY(h, cudaMemcpyAsync(B.confs[par], confs,
                                (size_t) n * 2 * N * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(B.energy[par], energy, n * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(B.weight[par], weight, n * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    if (slot_energy)
        CUDA_TRY(h, cudaMemcpyAsync(B.slot_energy, slot_energy,
                                    B.cap * sizeof(double),
                                    cudaMemcpyHostToDevice, h->stream));
    else
        CUDA_TRY(h, cudaMemcpyAsync(B.slot_energy, energy, n * sizeof(double),
                                    cudaMemcpyHostToDevice, h->stream));
    DmcCtl ctl{};
    ctl.step = sc->step;
    ctl.capacity_hits = sc->capacity_hits;
    ctl.W_prev = (int) n;
    ctl.W = (int) n;
    ctl.eref[0] = ctl.eref[1] = sc->ref_energy;
    ctl.tot_e = sc->total_energy;
    ctl.tot_w = sc->total_weight;
    ctl.last_energy = sc->energy;
    ctl.last_weight = sc->weight;
    ctl.last_accum = sc->accum_energy;
    ctl.W_global = sc->weight;
    ctl.pos_base[0] = ctl.pos_base[1] = global_slot_offset;
    CUDA_TRY(h, cudaMemcpyAsync(B.ctl, &ctl, sizeof ctl,
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->step_host = sc->step;
    h->n_ini = (int) n;
    h->dmc_ready = true;
    return QMCB_OK;
}

int qmcb_dmc_run_block(qmcb_handle *h, int64_t nts, int32_t eval_estimators,
                       double *energy, double *weight, uint64_t *num_walkers,
                       double *ref_energy, double *accum_energy,
                       double *density, double *ssf)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!h->dmc_ready) FAIL(h, QMCB_ERR_STATE, "qmcb_dmc_init not called");
    if (nts < 1) FAIL(h, QMCB_ERR_INVALID, "nts must be >= 1");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_log(h, nts);
    if (rc) return rc;
    const int M3 = 3 * h->dp.ssf_num_modes, NB = h->dp.density_num_bins;
    const bool do_ssf = eval_estimators && M3 > 0;
    const bool do_den = eval_estimators && NB > 0;
    const size_t cap_sz = (size_t) h->B.cap;
    if (do_ssf || do_den) {
        // qmc_base/dmc.py:901-909: every block starts from zeroed tables
        rc = ensure_est_log(h, nts);
        if (rc) return rc;
        if (do_ssf) {
            for (int i = 0; i < 2; ++i)
                CUDA_TRY(h, cudaMemsetAsync(h->ssf_aux[i], 0,
                                            cap_sz * M3 * sizeof(double),
                                            h->stream));
        }
        if (do_den) {
            for (int i = 0; i < 2; ++i)
                if (h->den_hist[i])
                    CUDA_TRY(h, cudaMemsetAsync(h->den_hist[i], 0,
                                                cap_sz * NB * sizeof(double),
                                                h->stream));
            CUDA_TRY(h, cudaMemsetAsync(h->den_total, 0,
                                        3 * (size_t) NB * sizeof(double),
                                        h->stream));
            CUDA_TRY(h, cudaMemsetAsync(h->den_hi, 0, sizeof(int),
                                        h->stream));
        }
    }
    long long est_launches = 0;
    DmcBufs &B = h->B;
    DmcLog L = h->L;
    const GroupGeom &g = h->geom;
    const int step_grid = (B.cap + g.G - 1) / g.G;
    if (h->comm) {
        rc = multi_setup(h);
        if (rc) return rc;
    }
    const DmcMulti X = h->X;
    const bool glob_a = h->comm && h->C.defer_weight;
    h->rebalanced_in_block = 0;
    if (h->profile_steps) {
        while ((long long) h->step_ev.size() < 2 * nts) {
            cudaEvent_t ev;
            CUDA_TRY(h, cudaEventCreate(&ev));
            h->step_ev.push_back(ev);
        }
    }
    // A block on one rank with no estimators and no per-launch events is a
    // fixed launch sequence with constant arguments: captured once into a
    // CUDA graph and replayed (small populations are launch-bound: 3 kernels
    // of a few microseconds each per time step).
    static const bool no_graph = getenv("QMCB_NO_GRAPH") != nullptr;
    const bool use_graph = !no_graph && !h->comm && !h->profile_steps
                           && !do_ssf && !do_den;
    const bool step_fast = step_fast_ok(h);
    // programmatic dependent launch between the kernels of a time step (see
    // pdl_enter); one rank only: with a communicator the NCCL kernels and
    // the second stream sit between them
    static const bool no_pdl = getenv("QMCB_NO_PDL") != nullptr;
    const bool pdl = !no_pdl && !h->comm;
    auto enqueue_step = [&](int64_t i) -> int {
        const int fuse = (glob_a && h->weights_pending) ? 1 : 0;
        CUDA_TRY(h, launch_pdl(branch_count_kernel, dim3(B.nblk),
                               dim3(BR_THREADS), 0, h->stream, pdl, B, h->C,
                               X, fuse));
        if (fuse) {
            // the previous step's update of the global array: after its
            // values have been read above, off the critical path
            CUDA_TRY(h, cudaEventRecord(h->ev_weighted, h->stream));
            CUDA_TRY(h, cudaStreamWaitEvent(h->pc_stream, h->ev_weighted, 0));
            multi_apply_kernel<<<h->sm_count * 4, 256, 0, h->pc_stream>>>(X);
        }
        CUDA_TRY(h, launch_pdl(branch_fill_kernel, dim3(B.nblk),
                               dim3(BR_THREADS), 0, h->stream, pdl, B, h->C,
                               L, h->comm ? 0 : 1));
        if (h->comm) {
            // Population control needs the GLOBAL {sum E, W}
            // (qmc_base/dmc.py:758-771) and, with the reference's stale-slot
            // weights, every rank needs the values the others write into
            // the global per-position array: one all-reduce of R + 2 doubles
            // (sums and the counts of all ranks) and one all-gather of 8
            // bytes per slot.  The step kernel of this time step needs none
            // of it (it reads E_ref of the previous step and its step index
            // from ctl->tcur), so all of this runs on a second stream next
            // to it; the weights (multi_weight_kernel) and the next step's
            // branching wait for it.
            CUDA_TRY(h, cudaEventRecord(h->ev_branched, h->stream));
            CUDA_TRY(h, cudaStreamWaitEvent(h->pc_stream, h->ev_branched, 0));
            multi_pack_kernel<<<glob_a ? (B.cap + 255) / 256 : 1, 256, 0,
                                h->pc_stream>>>(B, X, glob_a ? 1 : 0);
            NCCL_TRY(h, nccl_api()->AllReduce(X.vred, X.vred, X.R + 2,
                                              ncclDouble, ncclSum, h->comm,
                                              h->pc_stream));
            if (glob_a)
                NCCL_TRY(h, nccl_api()->AllGather(X.slab, X.gathered, B.cap,
                                                  ncclDouble, h->comm,
                                                  h->pc_stream));
            multi_finalize_kernel<<<1, 32, 0, h->pc_stream>>>(B, h->C, L, X);
            CUDA_TRY(h, cudaEventRecord(h->ev_controlled, h->pc_stream));
        }
        if (h->profile_steps)
            CUDA_TRY(h, cudaEventRecord(h->step_ev[2 * i], h->stream));
        if (step_fast)
            CUDA_TRY(h, launch_pdl(dmc_step_kernel<true>, dim3(step_grid),
                                   dim3(g.nthreads), (size_t) g.smem_bytes,
                                   h->stream, pdl, h->M, g, B, h->C));
        else
            CUDA_TRY(h, launch_pdl(dmc_step_kernel<false>, dim3(step_grid),
                                   dim3(g.nthreads), (size_t) g.smem_bytes,
                                   h->stream, pdl, h->M, g, B, h->C));
        if (h->profile_steps)
            CUDA_TRY(h, cudaEventRecord(h->step_ev[2 * i + 1], h->stream));
        if (h->comm) {
            CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_controlled, 0));
            // the weights of this step's children are formed by the next
            // branching (branch_count_kernel, fuse_weight) or, at the end of
            // the block / before a rebalance, by materialize_weights
            if (glob_a) h->weights_pending = true;
        }
        return QMCB_OK;
    };
    if (use_graph && (!h->block_graph || h->block_graph_nts != nts)) {
        drop_block_graph(h);
        cudaGraph_t graph = nullptr;
        CUDA_TRY(h, cudaStreamBeginCapture(h->stream,
                                           cudaStreamCaptureModeThreadLocal));
        dmc_block_begin_kernel<<<1, 32, 0, h->stream>>>(B);
        for (int64_t i = 0; i < nts; ++i) enqueue_step(i);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (ce != cudaSuccess || !graph) {
            cudaGetLastError();
            FAIL(h, QMCB_ERR_CUDA, std::string("graph capture of a DMC block "
                                               "failed: ")
                                       + cudaGetErrorString(ce));
        }
        ce = cudaGraphInstantiate(&h->block_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
            h->block_graph = nullptr;
            FAIL(h, QMCB_ERR_CUDA, std::string("cudaGraphInstantiate: ")
                                       + cudaGetErrorString(ce));
        }
        h->block_graph_nts = nts;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    if (use_graph) {
        CUDA_TRY(h, cudaGraphLaunch(h->block_graph, h->stream));
    } else {
        dmc_block_begin_kernel<<<1, 32, 0, h->stream>>>(B);
    }
    for (int64_t i = 0; i < nts && !use_graph; ++i) {
        rc = enqueue_step(i);
        if (rc) return rc;
        if (do_den) {
            rc = launch_density_step(h, i);
            if (rc) return rc;
            est_launches += 3;
        }
        if (do_ssf) {
            rc = launch_ssf_step(h, i);
            if (rc) return rc;
            est_launches += 3;
        }
        // Sharded runs without per-slot estimator state (the forward-walking
        // rows of pure S(k), the bin lists of the density): every
        // `rebalance_every` steps look at the counts of all ranks (one
        // 8 (R + 1)-byte read) and even the slabs out when they have drifted
        // apart by more than 1 %, or when one of them nears its capacity
        // (the reference truncates at the GLOBAL capacity only).
        if (h->comm && !do_den && (!do_ssf || !h->dp.ssf_as_pure)
            && h->rebalance_every > 0
            && (i + 1) % h->rebalance_every == 0 && i + 1 < nts) {
            std::vector<long long> offs(X.R + 1);
            CUDA_TRY(h, cudaMemcpyAsync(offs.data(), X.offs,
                                        (X.R + 1) * sizeof(long long),
                                        cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            long long mx = 0, mn = 1ll << 62;
            for (int q = 0; q < X.R; ++q) {
                mx = std::max(mx, offs[q + 1] - offs[q]);
                mn = std::min(mn, offs[q + 1] - offs[q]);
            }
            if (mx - mn > 1 && ((double) mx > 1.01 * (double) mn
                                || (double) mx > 0.9 * (double) B.cap)) {
                rc = materialize_weights(h);
                if (rc) return rc;
                int64_t mv = 0;
                rc = qmcb_dmc_rebalance(h, &mv);
                if (rc) return rc;
                h->rebalanced_in_block += mv;
            }
        }
    }
    if (h->comm) {
        // the weights of the last step and the control stream's last update
        // of the global array are part of this block
        rc = materialize_weights(h);
        if (rc) return rc;
        CUDA_TRY(h, cudaEventRecord(h->ev_controlled, h->pc_stream));
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_controlled, 0));
    }
    if (h->comm && (do_den || do_ssf)) {
        // the tables hold this rank's partial sums: global sums on the
        // device, before they leave for the host
        if (do_den)
            NCCL_TRY(h, nccl_api()->AllReduce(h->den_iter, h->den_iter,
                                              nts * (size_t) NB, ncclDouble,
                                              ncclSum, h->comm, h->stream));
        if (do_ssf)
            NCCL_TRY(h, nccl_api()->AllReduce(h->ssf_iter, h->ssf_iter,
                                              nts * (size_t) M3, ncclDouble,
                                              ncclSum, h->comm, h->stream));
    }
    if (h->rb.K > 0) {
        // same expression as on_the_fly_obj_data_order (stats/reblock.py:447)
        const int order = (int) std::floor(std::log((double) nts)
                                           / std::log(2.0));
        reblock_series_kernel<<<1, 32, 0, h->stream>>>(L, nts, order, h->rb);
    }
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    h->step_host += nts;
    h->last_launches = 1 + (3 + (h->comm ? (glob_a ? 3 : 2) : 0)) * nts
                       + (glob_a ? 1 : 0) + est_launches;
    if (density && do_den)
        CUDA_TRY(h, cudaMemcpyAsync(density, h->den_iter,
                                    nts * (size_t) NB * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (ssf && do_ssf)
        CUDA_TRY(h, cudaMemcpyAsync(ssf, h->ssf_iter,
                                    nts * (size_t) M3 * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (energy)
        CUDA_TRY(h, cudaMemcpyAsync(energy, L.energy, nts * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (weight)
        CUDA_TRY(h, cudaMemcpyAsync(weight, L.weight, nts * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (num_walkers)
        CUDA_TRY(h, cudaMemcpyAsync(num_walkers, L.num_walkers,
                                    nts * sizeof(uint64_t),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (ref_energy)
        CUDA_TRY(h, cudaMemcpyAsync(ref_energy, L.ref_energy,
                                    nts * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (accum_energy)
        CUDA_TRY(h, cudaMemcpyAsync(accum_energy, L.accum_energy,
                                    nts * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_total_ms = ms;
    h->last_step_ms = 0.0;
    if (h->profile_steps) {
        double acc = 0.0;
        for (int64_t i = 0; i < nts; ++i) {
            float sms = 0.f;
            CUDA_TRY(h, cudaEventElapsedTime(&sms, h->step_ev[2 * i],
                                             h->step_ev[2 * i + 1]));
            acc += sms;
        }
        h->last_step_ms = acc;
    }
    return QMCB_OK;
}

int qmcb_dmc_reblock_reset(qmcb_handle *h, int32_t max_order)
{
    if (!h) return QMCB_ERR_INVALID;
    if (max_order >= RB_MAX_ORDERS)
        FAIL(h, QMCB_ERR_INVALID, "max_order too large");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->rb.sum); cudaFree(h->rb.sqr); cudaFree(h->rb.nblk);
    h->rb = ReblockTables{};
    if (max_order < 0) return QMCB_OK;          // switched off
    const int K = max_order + 1;
    const size_t n = (size_t) RB_COLS * K;
    CUDA_TRY(h, cudaMalloc(&h->rb.sum, n * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&h->rb.sqr, n * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&h->rb.nblk, n * sizeof(long long)));
    CUDA_TRY(h, cudaMemsetAsync(h->rb.sum, 0, n * sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->rb.sqr, 0, n * sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->rb.nblk, 0, n * sizeof(long long),
                                h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->rb.K = K;
    return QMCB_OK;
}

int qmcb_dmc_reblock_get(qmcb_handle *h, double *means_sum,
                         double *means_sqr_sum, int64_t *num_blocks)
{
    if (!h) return QMCB_ERR_INVALID;
    if (h->rb.K < 1)
        FAIL(h, QMCB_ERR_STATE, "qmcb_dmc_reblock_reset not called");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t) RB_COLS * h->rb.K;
    if (means_sum)
        CUDA_TRY(h, cudaMemcpyAsync(means_sum, h->rb.sum, n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (means_sqr_sum)
        CUDA_TRY(h, cudaMemcpyAsync(means_sqr_sum, h->rb.sqr,
                                    n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (num_blocks)
        CUDA_TRY(h, cudaMemcpyAsync(num_blocks, h->rb.nblk,
                                    n * sizeof(long long),
                                    cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

int qmcb_set_profiling(qmcb_handle *h, int32_t on)
{
    if (!h) return QMCB_ERR_INVALID;
    h->profile_steps = on != 0;
    return QMCB_OK;
}

int qmcb_last_block_stats(qmcb_handle *h, double *total_ms,
                          double *step_kernel_ms, int64_t *launches)
{
    if (!h) return QMCB_ERR_INVALID;
    if (total_ms) *total_ms = h->last_total_ms;
    if (step_kernel_ms) *step_kernel_ms = h->last_step_ms;
    if (launches) *launches = h->last_launches;
    return QMCB_OK;
}

int qmcb_dmc_get_state(qmcb_handle *h, double *confs, double *energy,
                       double *weight, uint8_t *mask, int64_t *cloning_ref,
                       qmcb_state_scalars *scalars)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!h->dmc_ready) FAIL(h, QMCB_ERR_STATE, "qmcb_dmc_init not called");
    CUDA_TRY(h, cudaSetDevice(h->device));
    DmcCtl ctl;
    int rc = read_ctl(h, ctl);
    if (rc) return rc;
    fill_scalars(h, ctl, scalars);
    DmcBufs &B = h->B;
    const int N = h->M.nop;
    const long long cap = B.cap;
    bool identity = ctl.step == 0;
    // the last executed step `step-1` branched from buffer (step-1)&1
    int par = identity ? 0 : (int) ((ctl.step - 1) & 1);
    int W = identity ? ctl.W_prev : ctl.W;
    size_t nc = (size_t) cap * 2 * N * sizeof(double);
    size_t total = nc + 2 * cap * sizeof(double) + cap * sizeof(long long)
                   + cap;
    rc = ensure_scratch(h, total);
    if (rc) return rc;
    double *d_c = h->d_scratch;
    double *d_e = d_c + (size_t) cap * 2 * N;
    double *d_w = d_e + cap;
    long long *d_r = reinterpret_cast<long long *>(d_w + cap);
    unsigned char *d_m = reinterpret_cast<unsigned char *>(d_r + cap);
    gather_state_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(
        B.confs[par], B.energy[par], B.ref, W, B.cap, N, identity,
        confs ? d_c : nullptr, d_e, d_w, d_m, d_r);
    CUDA_TRY(h, cudaGetLastError());
    if (confs)
        CUDA_TRY(h, cudaMemcpyAsync(confs, d_c, nc, cudaMemcpyDeviceToHost,
                                    h->stream));
    if (energy)
        CUDA_TRY(h, cudaMemcpyAsync(energy, d_e, cap * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (weight)
        CUDA_TRY(h, cudaMemcpyAsync(weight, d_w, cap * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (mask)
        CUDA_TRY(h, cudaMemcpyAsync(mask, d_m, cap, cudaMemcpyDeviceToHost,
                                    h->stream));
    if (cloning_ref)
        CUDA_TRY(h, cudaMemcpyAsync(cloning_ref, d_r, cap * sizeof(long long),
                                    cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

int qmcb_dmc_get_next(qmcb_handle *h, double *confs, double *energy,
                      double *weight, double *slot_energy,
                      qmcb_state_scalars *scalars)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!h->dmc_ready) FAIL(h, QMCB_ERR_STATE, "qmcb_dmc_init not called");
    CUDA_TRY(h, cudaSetDevice(h->device));
    DmcCtl ctl;
    int rc = read_ctl(h, ctl);
    if (rc) return rc;
    fill_scalars(h, ctl, scalars);
    if (scalars) scalars->num_walkers = ctl.W_prev;
    DmcBufs &B = h->B;
    const int N = h->M.nop;
    int par = (int) (ctl.step & 1);
    size_t n = (size_t) ctl.W_prev;
    if (confs)
        CUDA_TRY(h, cudaMemcpyAsync(confs, B.confs[par],
                                    n * 2 * N * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (energy)
        CUDA_TRY(h, cudaMemcpyAsync(energy, B.energy[par], n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (weight)
        CUDA_TRY(h, cudaMemcpyAsync(weight, B.weight[par], n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (slot_energy) {
        if (h->comm && h->C.defer_weight && h->multi_ready) {
            // the local view of the global per-position array (the import
            // convention of multi_setup)
            multi_export_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(
                h->X, ctl.pos_base[ctl.step & 1], B.slot_energy);
            CUDA_TRY(h, cudaGetLastError());
        }
        CUDA_TRY(h, cudaMemcpyAsync(slot_energy, B.slot_energy,
                                    B.cap * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

int qmcb_measure_fp64_peak(int device, double *tflops, double *ms_out)
{
    if (!tflops) return QMCB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return QMCB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return QMCB_ERR_CUDA;
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    double *d = nullptr;
    if (cudaMalloc(&d, (size_t) grid * 256 * sizeof(double)) != cudaSuccess)
        return QMCB_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<grid, 256>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            cudaFree(d);
            return QMCB_ERR_CUDA;
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    double flops = 2.0 * 64.0 * iters * (double) grid * 256.0;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return QMCB_OK;
}

int qmcb_measure_fp64_sustained(int device, double seconds, double *tflops)
{
    if (!tflops || !(seconds > 0)) return QMCB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return QMCB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return QMCB_ERR_CUDA;
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    double *d = nullptr;
    if (cudaMalloc(&d, (size_t) grid * 256 * sizeof(double)) != cudaSuccess)
        return QMCB_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // one launch to learn the duration, then a fixed count back to back
    cudaEventRecord(e0);
    fp64_peak_kernel<<<grid, 256>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms1 = 0.f;
    cudaEventElapsedTime(&ms1, e0, e1);
    int n = (int) (seconds * 1e3 / (ms1 > 0.01f ? ms1 : 0.01f));
    if (n < 4) n = 4;
    if (n > 100000) n = 100000;
    cudaEventRecord(e0);
    for (int i = 0; i < n; ++i)
        fp64_peak_kernel<<<grid, 256>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess || !(ms > 0)) return QMCB_ERR_CUDA;
    double flops = 2.0 * 64.0 * iters * (double) grid * 256.0 * n;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return QMCB_OK;
}

void *qmcb_stream(qmcb_handle *h) { return h ? (void *) h->stream : nullptr; }

void *qmcb_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t) bytes, cudaHostAllocDefault) != cudaSuccess)
        return nullptr;
    return p;
}

void qmcb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int qmcb_comm_unique_id(uint8_t *id)
{
    NcclApi *api = nccl_api();
    if (!api || !id) return QMCB_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId uid;
    if (api->GetUniqueId(&uid) != ncclSuccess) return QMCB_ERR_NCCL;
    memcpy(id, &uid, sizeof uid);
    return QMCB_OK;
}

int qmcb_comm_init(qmcb_handle *h, const uint8_t *id, int32_t world_size,
                   int32_t rank)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!id || world_size < 1 || rank < 0 || rank >= world_size)
        FAIL(h, QMCB_ERR_INVALID, "bad communicator arguments");
    NcclApi *api = nccl_api();
    if (!api) FAIL(h, QMCB_ERR_NCCL, "NCCL not loadable: libnccl.so.2");
    if (h->comm) FAIL(h, QMCB_ERR_STATE, "communicator already initialised");
    CUDA_TRY(h, cudaSetDevice(h->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    // The per-step collectives are a few KB to ~1 MB and run NEXT TO the step
    // kernel, which fills every SM: keep NCCL's kernels small (few CTAs) so
    // that they fit into the resources one retiring step CTA frees.
    bool made = false;
    if (api->CommInitRankConfig && !getenv("QMCB_NCCL_DEFAULT_CONFIG")) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = 4;
        if (const char *e = getenv("QMCB_NCCL_MAX_CTAS"))
            cfg.maxCTAs = std::max(1, atoi(e));
        ncclResult_t r = api->CommInitRankConfig(&h->comm, world_size, uid,
                                                 rank, &cfg);
        made = r == ncclSuccess;
        if (!made) h->comm = nullptr;
    }
    if (!made)
        NCCL_TRY(h, api->CommInitRank(&h->comm, world_size, uid, rank));
    h->world = world_size;
    h->rank = rank;
    if (h->dmc_ready) {
        // a population loaded before the communicator existed: switch to
        // the sharded weight path now
        h->C.defer_weight = h->dp.energy_mode == 0 ? 1 : 0;
        h->multi_ready = false;
        drop_block_graph(h);
    }
    if (const char *e = getenv("QMCB_REBALANCE_EVERY"))
        h->rebalance_every = atoll(e);
    CUDA_TRY(h, cudaMalloc(&h->d_counts, world_size * sizeof(long long)));
    // NCCL sets its peer-to-peer channels up lazily, on the first send/recv
    // of every (source, destination) pair, and that costs tens of
    // milliseconds per pair.  The rebalance may pair any two ranks, so open
    // all channels (and the all-reduce / all-gather rings) now rather than
    // inside a sampling block.
    CUDA_TRY(h, cudaMemsetAsync(h->d_counts, 0,
                                world_size * sizeof(long long), h->stream));
    long long *d_tmp = nullptr;
    CUDA_TRY(h, cudaMalloc(&d_tmp, 2 * world_size * sizeof(long long)));
    CUDA_TRY(h, cudaMemsetAsync(d_tmp, 0, 2 * world_size * sizeof(long long),
                                h->stream));
    auto warm = [&]() -> int {
        for (int p = 0; p < world_size; ++p) {
            if (p == rank) continue;
            NCCL_TRY(h, api->Send(h->d_counts + rank, 1, ncclInt64, p,
                                  h->comm, h->stream));
            NCCL_TRY(h, api->Recv(d_tmp + p, 1, ncclInt64, p, h->comm,
                                  h->stream));
        }
        return QMCB_OK;
    };
    NCCL_TRY(h, api->GroupStart());
    {
        int wrc = warm();
        ncclResult_t ge = api->GroupEnd();
        if (wrc) return wrc;
        NCCL_TRY(h, ge);
    }
    NCCL_TRY(h, api->AllReduce(d_tmp + world_size, d_tmp + world_size, 2,
                               ncclDouble, ncclSum, h->comm, h->stream));
    NCCL_TRY(h, api->AllGather(h->d_counts + rank, h->d_counts, 1, ncclInt64,
                               h->comm, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaFree(d_tmp));
    // highest priority: the block scheduler places the (small) collective
    // and population-control kernels ahead of the step kernel's queue of
    // CTAs instead of after it
    int prio_lo = 0, prio_hi = 0;
    CUDA_TRY(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUDA_TRY(h, cudaStreamCreateWithPriority(&h->pc_stream,
                                             cudaStreamNonBlocking, prio_hi));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_weighted,
                                         cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_branched,
                                         cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_controlled,
                                         cudaEventDisableTiming));
    return QMCB_OK;
}

// The exchange plan of the order-preserving rebalance, host arithmetic only
// (exported so that the partition logic can be tested without GPUs).
// counts[r] = live walkers of rank r.  For peer p:
//   send[2p], send[2p+1] = (offset in my current slab, n) of walkers p takes
//   recv[2p], recv[2p+1] = (offset in my NEW slab, n) of walkers p gives me
// (p == rank describes the part that stays).  new_count = my new population.
int qmcb_rebalance_plan(const int64_t *counts, int32_t world, int32_t rank,
                        int64_t *send, int64_t *recv, int64_t *new_count)
{
    if (!counts || world < 1 || rank < 0 || rank >= world || !send || !recv)
        return QMCB_ERR_INVALID;
    std::vector<long long> cur0(world + 1, 0), new0(world + 1, 0);
    for (int r = 0; r < world; ++r) {
        if (counts[r] < 0) return QMCB_ERR_INVALID;
        cur0[r + 1] = cur0[r] + counts[r];
    }
    const long long T = cur0[world];
    for (int r = 0; r <= world; ++r) new0[r] = T * r / world;
    auto overlap = [](long long a0, long long a1, long long b0, long long b1,
                      long long &lo, long long &hi) {
        lo = a0 > b0 ? a0 : b0;
        hi = a1 < b1 ? a1 : b1;
        return hi > lo;
    };
    for (int p = 0; p < world; ++p) {
        long long lo, hi;
        send[2 * p] = send[2 * p + 1] = 0;
        recv[2 * p] = recv[2 * p + 1] = 0;
        if (overlap(cur0[rank], cur0[rank + 1], new0[p], new0[p + 1], lo,
                    hi)) {
            send[2 * p] = lo - cur0[rank];
            send[2 * p + 1] = hi - lo;
        }
        if (overlap(cur0[p], cur0[p + 1], new0[rank], new0[rank + 1], lo,
                    hi)) {
            recv[2 * p] = lo - new0[rank];
            recv[2 * p + 1] = hi - lo;
        }
    }
    if (new_count) *new_count = new0[rank + 1] - new0[rank];
    return QMCB_OK;
}

// Order-preserving rebalance.  The global ensemble is the
// concatenation of the ranks' live walkers; the new partition gives rank r
// the global range [T r / R, T (r + 1) / R).  Every rank sends the parts of
// its current range that fall into other ranks' new ranges and receives the
// parts of its new range that others hold -- with fluctuating populations
// these are thin slices at the boundaries with the ring neighbours.  Works on
// the evolved population the next step will branch from.
int qmcb_dmc_rebalance(qmcb_handle *h, int64_t *moved)
{
    if (!h) return QMCB_ERR_INVALID;
    if (moved) *moved = 0;
    if (!h->dmc_ready) FAIL(h, QMCB_ERR_STATE, "qmcb_dmc_init not called");
    if (!h->comm || h->world == 1) return QMCB_OK;
    NcclApi *api = nccl_api();
    CUDA_TRY(h, cudaSetDevice(h->device));
    DmcBufs &B = h->B;
    const int R = h->world, me = h->rank, N = h->M.nop;
    int rc = materialize_weights(h);    // the weights travel with the walkers
    if (rc) return rc;
    DmcCtl ctl;
    rc = read_ctl(h, ctl);
    if (rc) return rc;
    const int par = (int) (ctl.step & 1);
    long long mine = ctl.W_prev;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_counts + me, &mine, sizeof mine,
                                cudaMemcpyHostToDevice, h->stream));
    NCCL_TRY(h, api->AllGather(h->d_counts + me, h->d_counts, 1, ncclInt64,
                               h->comm, h->stream));
    std::vector<long long> cnt(R);
    CUDA_TRY(h, cudaMemcpyAsync(cnt.data(), h->d_counts,
                                R * sizeof(long long), cudaMemcpyDeviceToHost,
                                h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    std::vector<int64_t> cnt64(cnt.begin(), cnt.end());
    std::vector<int64_t> plan_send(2 * R), plan_recv(2 * R);
    int64_t new_n64 = 0;
    qmcb_rebalance_plan(cnt64.data(), R, me, plan_send.data(),
                        plan_recv.data(), &new_n64);
    const long long new_n = new_n64;
    if (new_n > B.cap) FAIL(h, QMCB_ERR_STATE, "rebalance exceeds capacity");
    bool any = false;
    for (int p = 0; p < R; ++p)
        any = any || (p != me && (plan_send[2 * p + 1] || plan_recv[2 * p + 1]));
    // sends and receives pair up rank by rank: a rank whose slab keeps its
    // walkers has nothing to post
    if (!any) return QMCB_OK;

    // staging: [new_n] x (confs, energy, weight)
    const size_t row = (size_t) 2 * N;
    size_t need = (size_t) new_n * (row + 2) * sizeof(double);
    rc = ensure_scratch(h, need);
    if (rc) return rc;
    double *s_confs = h->d_scratch;
    double *s_energy = s_confs + (size_t) new_n * row;
    double *s_weight = s_energy + new_n;
    long long sent = 0;
    // every call between GroupStart and GroupEnd is checked, but a failure
    // must not leave the group open (the peers would hang in theirs)
    auto exchange = [&]() -> int {
        for (int p = 0; p < R; ++p) {
            // my current walkers that p will own
            if (plan_send[2 * p + 1] > 0) {
                long long src = plan_send[2 * p], n = plan_send[2 * p + 1];
                if (p == me) {
                    long long dst = plan_recv[2 * p];
                    CUDA_TRY(h, cudaMemcpyAsync(
                                    s_confs + dst * row,
                                    B.confs[par] + src * row,
                                    n * row * sizeof(double),
                                    cudaMemcpyDeviceToDevice, h->stream));
                    CUDA_TRY(h, cudaMemcpyAsync(s_energy + dst,
                                                B.energy[par] + src,
                                                n * sizeof(double),
                                                cudaMemcpyDeviceToDevice,
                                                h->stream));
                    CUDA_TRY(h, cudaMemcpyAsync(s_weight + dst,
                                                B.weight[par] + src,
                                                n * sizeof(double),
                                                cudaMemcpyDeviceToDevice,
                                                h->stream));
                } else {
                    NCCL_TRY(h, api->Send(B.confs[par] + src * row, n * row,
                                          ncclDouble, p, h->comm, h->stream));
                    NCCL_TRY(h, api->Send(B.energy[par] + src, n, ncclDouble,
                                          p, h->comm, h->stream));
                    NCCL_TRY(h, api->Send(B.weight[par] + src, n, ncclDouble,
                                          p, h->comm, h->stream));
                    sent += n;
                }
            }
            // walkers of p that I will own
            if (p != me && plan_recv[2 * p + 1] > 0) {
                long long dst = plan_recv[2 * p], n = plan_recv[2 * p + 1];
                NCCL_TRY(h, api->Recv(s_confs + dst * row, n * row,
                                      ncclDouble, p, h->comm, h->stream));
                NCCL_TRY(h, api->Recv(s_energy + dst, n, ncclDouble, p,
                                      h->comm, h->stream));
                NCCL_TRY(h, api->Recv(s_weight + dst, n, ncclDouble, p,
                                      h->comm, h->stream));
            }
        }
        return QMCB_OK;
    };
    NCCL_TRY(h, api->GroupStart());
    rc = exchange();
    {
        ncclResult_t ge = api->GroupEnd();
        if (rc) return rc;
        NCCL_TRY(h, ge);
    }
    CUDA_TRY(h, cudaMemcpyAsync(B.confs[par], s_confs,
                                new_n * row * sizeof(double),
                                cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(B.energy[par], s_energy,
                                new_n * sizeof(double),
                                cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(B.weight[par], s_weight,
                                new_n * sizeof(double),
                                cudaMemcpyDeviceToDevice, h->stream));
    int w = (int) new_n;
    CUDA_TRY(h, cudaMemcpyAsync(&B.ctl->W_prev, &w, sizeof w,
                                cudaMemcpyHostToDevice, h->stream));
    // the walkers keep their global positions (the exchange preserves the
    // order); this rank's slab now starts at another one
    long long total = 0, base = 0;
    for (int p = 0; p < R; ++p) total += cnt[p];
    base = total * me / R;
    CUDA_TRY(h, cudaMemcpyAsync(&B.ctl->pos_base[par], &base, sizeof base,
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (moved) *moved = sent;
    return QMCB_OK;
}

int qmcb_vmc_init(qmcb_handle *h, const qmcb_vmc_params *params,
                  const double *confs, int64_t num_chains)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!params || !confs || num_chains < 1)
        FAIL(h, QMCB_ERR_INVALID, "bad arguments");
    if (!(params->upper_bound > params->lower_bound))
        FAIL(h, QMCB_ERR_INVALID, "upper_bound must exceed lower_bound");
    if (params->ssf_num_modes < 0 || params->ssf_num_modes > 65536)
        FAIL(h, QMCB_ERR_INVALID, "bad ssf_num_modes");
    if (params->proposal != 0 && params->proposal != 1)
        FAIL(h, QMCB_ERR_INVALID, "proposal must be 0 (uniform) or 1 "
                                  "(gaussian)");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = h->M.nop, M = params->ssf_num_modes;
    size_t vsm = (size_t) h->geom.smem_bytes;
    if (vsm > (size_t) h->max_smem)
        FAIL(h, QMCB_ERR_INVALID, "boson_number too large for the VMC kernel");
    CUDA_TRY(h, cudaFuncSetAttribute(
                    vmc_block_kernel<false>,
                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) vsm));
    CUDA_TRY(h, cudaFuncSetAttribute(
                    vmc_block_kernel<true>,
                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) vsm));
    const size_t C = (size_t) num_chains;
    // a re-initialisation of the same shape (chains handed back from the
    // host block after block) keeps its device buffers
    const bool same_shape = h->vmc_ready && h->vmc_chains == num_chains
                            && h->vp.ssf_num_modes == M;
    if (!same_shape) {
        free_vmc(h);
        CUDA_TRY(h, cudaMalloc(&h->V.confs, C * 2 * N * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&h->V.lnpsi, C * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&h->V.eprev, C * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&h->vmc_sum_e, C * 2 * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&h->vmc_acc, C * sizeof(double)));
        if (M) {
            CUDA_TRY(h, cudaMalloc(&h->V.ssfprev,
                                   C * M * 3 * sizeof(double)));
            CUDA_TRY(h, cudaMalloc(&h->vmc_sum_ssf,
                                   C * M * 3 * sizeof(double)));
        }
    }
    h->vmc_ready = false;
    h->vp = *params;
    if (M)
        CUDA_TRY(h, cudaMemsetAsync(h->V.ssfprev, 0,
                                    C * M * 3 * sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->V.eprev, 0, C * sizeof(double),
                                h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->V.confs, confs,
                                C * 2 * N * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    // Sampling.build_state (mrbp_qmc/vmc.py:145-170): ln|Psi| of the
    // initial configurations
    EvalArgs a{};
    a.confs = h->V.confs; a.nconf = num_chains; a.lnpsi = h->V.lnpsi;
    int rc = launch_model_eval(h, a, true, false);
    if (rc) return rc;
    // the table path of the block kernel needs every position in [0, L]
    h->vmc_fast = h->M.tt.zt != nullptr && params->lower_bound == 0.0
                  && params->upper_bound - params->lower_bound == h->M.L;
    int outside = 0;
    if (h->vmc_fast) {
        rc = ensure_scratch(h, 8);
        if (rc) return rc;
        int *d_flag = reinterpret_cast<int *>(h->d_scratch);
        CUDA_TRY(h, cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
        const long long total = (long long) C * N;
        const int grid = (int) std::min<long long>(
            (total + 255) / 256, (long long) h->sm_count * 16);
        positions_outside_kernel<<<grid, 256, 0, h->stream>>>(
            h->V.confs, (long long) C, N, h->M.L, d_flag);
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaMemcpyAsync(&outside, d_flag, sizeof(int),
                                    cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (outside) h->vmc_fast = false;
    h->vmc_chains = num_chains;
    h->vmc_gstep = 0;
    h->vmc_first = 1;
    h->vmc_ready = true;
    return QMCB_OK;
}

static int vmc_run_block_impl(qmcb_handle *h, int64_t ns, double *lnpsi,
                              double *energy, uint8_t *move_stat, double *ssf,
                              double *accept_rate, double *sum_energy,
                              double *sum_ssf, double *confs);

int qmcb_vmc_run_block(qmcb_handle *h, int64_t ns, double *lnpsi,
                       double *energy, uint8_t *move_stat, double *ssf,
                       double *accept_rate, double *sum_energy,
                       double *sum_ssf)
{
    return vmc_run_block_impl(h, ns, lnpsi, energy, move_stat, ssf,
                              accept_rate, sum_energy, sum_ssf, nullptr);
}

int qmcb_vmc_one_body_density(qmcb_handle *h, const double *offsets,
                              int32_t num_offsets, double *out)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!h->vmc_ready) FAIL(h, QMCB_ERR_STATE, "qmcb_vmc_init not called");
    if (num_offsets < 0 || (num_offsets > 0 && (!offsets || !out)))
        FAIL(h, QMCB_ERR_INVALID, "bad arguments");
    if (num_offsets == 0) return QMCB_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t C = (size_t) h->vmc_chains, S = (size_t) num_offsets;
    int rc = ensure_scratch(h, (S + C * S) * sizeof(double));
    if (rc) return rc;
    double *d_off = h->d_scratch, *d_out = d_off + S;
    CUDA_TRY(h, cudaMemcpyAsync(d_off, offsets, S * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    rc = qmcb_one_body_density_device(h, h->V.confs, h->vmc_chains, d_off,
                                      num_offsets, d_out);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(out, d_out, C * S * sizeof(double),
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

int qmcb_vmc_run_chain(qmcb_handle *h, int64_t ns, double *lnpsi,
                       double *energy, uint8_t *move_stat, double *confs,
                       double *accept_rate)
{
    if (h && !confs) FAIL(h, QMCB_ERR_INVALID, "confs must not be null");
    return vmc_run_block_impl(h, ns, lnpsi, energy, move_stat, nullptr,
                              accept_rate, nullptr, nullptr, confs);
}

static int vmc_run_block_impl(qmcb_handle *h, int64_t ns, double *lnpsi,
                              double *energy, uint8_t *move_stat, double *ssf,
                              double *accept_rate, double *sum_energy,
                              double *sum_ssf, double *confs)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!h->vmc_ready) FAIL(h, QMCB_ERR_STATE, "qmcb_vmc_init not called");
    if (ns < 1) FAIL(h, QMCB_ERR_INVALID, "ns must be >= 1");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t C = (size_t) h->vmc_chains;
    const int M = h->vp.ssf_num_modes;
    if (ssf && !M) FAIL(h, QMCB_ERR_INVALID, "S(k) estimator is off");
    // device staging for the optional per-step series
    size_t n_ser = C * (size_t) ns;
    size_t bytes = 0;
    size_t off_ln = bytes; if (lnpsi) bytes += n_ser * sizeof(double);
    size_t off_e = bytes; if (energy) bytes += n_ser * sizeof(double);
    size_t off_ssf = bytes; if (ssf) bytes += n_ser * M * 3 * sizeof(double);
    size_t off_st = bytes; if (move_stat) bytes += (n_ser + 7) / 8 * 8;
    const size_t conf_bytes = n_ser * 2 * (size_t) h->M.nop * sizeof(double);
    size_t off_cf = bytes; if (confs) bytes += conf_bytes;
    if (bytes > ((size_t) 64 << 30))
        FAIL(h, QMCB_ERR_INVALID, "per-step series of this block exceed "
                                  "64 GB: use fewer steps per call");
    int rc = ensure_scratch(h, bytes ? bytes : 8);
    if (rc) return rc;
    char *scr = reinterpret_cast<char *>(h->d_scratch);
    VmcArgs a{};
    a.nchains = h->vmc_chains; a.ns = ns;
    a.gstep0 = h->vmc_gstep; a.chain_offset = h->vp.chain_offset;
    a.first = h->vmc_first; a.M = M; a.seed = h->vp.rng_seed;
    a.proposal = h->vp.proposal;
    a.spread = h->vp.move_spread; a.z_min = h->vp.lower_bound;
    a.size = h->vp.upper_bound - h->vp.lower_bound;
    a.two_over_L = 2.0 / h->M.L;
    a.out_lnpsi = lnpsi ? reinterpret_cast<double *>(scr + off_ln) : nullptr;
    a.out_energy = energy ? reinterpret_cast<double *>(scr + off_e) : nullptr;
    a.out_ssf = ssf ? reinterpret_cast<double *>(scr + off_ssf) : nullptr;
    a.out_stat = move_stat ? reinterpret_cast<unsigned char *>(scr + off_st)
                           : nullptr;
    a.out_confs = confs ? reinterpret_cast<double *>(scr + off_cf) : nullptr;
    a.accept_rate = h->vmc_acc;
    a.sum_energy = h->vmc_sum_e;
    // the block sums of S(k) are always accumulated on the device (that is
    // the estimator); `sum_ssf` only says whether they are shipped
    a.sum_ssf = M ? h->vmc_sum_ssf : nullptr;
    if (a.sum_ssf)
        CUDA_TRY(h, cudaMemsetAsync(h->vmc_sum_ssf, 0,
                                    C * M * 3 * sizeof(double), h->stream));
    const GroupGeom &g = h->geom;
    size_t vsm = (size_t) g.smem_bytes;
    long long ctas = ((long long) C + g.G - 1) / g.G;
    int grid = (int) std::min<long long>(ctas, (long long) h->sm_count * 64);
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    if (h->vmc_fast)
        vmc_block_kernel<true><<<grid, g.nthreads, vsm, h->stream>>>(
            h->M, g, h->V, a);
    else
        vmc_block_kernel<false><<<grid, g.nthreads, vsm, h->stream>>>(
            h->M, g, h->V, a);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    h->vmc_gstep += ns - (h->vmc_first ? 1 : 0);
    h->vmc_first = 0;
    h->last_launches = 1;
    auto d2h = [&](void *dst, const void *src, size_t n) {
        return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost,
                               h->stream);
    };
    if (lnpsi) CUDA_TRY(h, d2h(lnpsi, a.out_lnpsi, n_ser * sizeof(double)));
    if (energy) CUDA_TRY(h, d2h(energy, a.out_energy, n_ser * sizeof(double)));
    if (ssf) CUDA_TRY(h, d2h(ssf, a.out_ssf, n_ser * M * 3 * sizeof(double)));
    if (move_stat) CUDA_TRY(h, d2h(move_stat, a.out_stat, n_ser));
    if (confs) CUDA_TRY(h, d2h(confs, a.out_confs, conf_bytes));
    if (accept_rate)
        CUDA_TRY(h, d2h(accept_rate, h->vmc_acc, C * sizeof(double)));
    if (sum_energy)
        CUDA_TRY(h, d2h(sum_energy, h->vmc_sum_e, C * 2 * sizeof(double)));
    if (sum_ssf && M)
        CUDA_TRY(h, d2h(sum_ssf, h->vmc_sum_ssf, C * M * 3 * sizeof(double)));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_total_ms = ms;
    h->last_step_ms = ms;
    return QMCB_OK;
}

int qmcb_vmc_get_state(qmcb_handle *h, double *confs, double *lnpsi)
{
    if (!h) return QMCB_ERR_INVALID;
    if (!h->vmc_ready) FAIL(h, QMCB_ERR_STATE, "qmcb_vmc_init not called");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t C = (size_t) h->vmc_chains;
    const int N = h->M.nop;
    if (confs)
        CUDA_TRY(h, cudaMemcpyAsync(confs, h->V.confs,
                                    C * 2 * N * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (lnpsi)
        CUDA_TRY(h, cudaMemcpyAsync(lnpsi, h->V.lnpsi, C * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return QMCB_OK;
}

}  // extern "C"
