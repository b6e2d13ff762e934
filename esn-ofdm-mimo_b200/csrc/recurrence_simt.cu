// C-ABI entry points of the SIMT reservoir recurrence; the kernels live in recurrence_simt.cuh
// and are instantiated per dtype in recurrence_simt_f32.cu / recurrence_simt_f64.cu.
#include "recurrence_simt.cuh"

int esn_simt_launch_f32(const esn_simt::RecParams &p, cudaStream_t st);
int esn_simt_launch_f64(const esn_simt::RecParams &p, cudaStream_t st);
int esn_cluster_launch(const esn_simt::RecParams &p, int dtype, cudaStream_t st);   // recurrence_cluster.cu

int esn_cluster_auto_limit(const esn_simt::RecParams &p, int dtype);                  // recurrence_cluster.cu
int esn_dmma_launch(const esn_simt::RecParams &p, cudaStream_t st);           // recurrence_dmma.cu
bool esn_dmma_enabled();

// Batches of at most this many frames take the cluster kernel (weights resident in the shared memory of a
// thread-block cluster, ~3 us per time step) instead of the streaming SIMT kernel (~90 us per step while its
// grid does not fill the GPU).  -1 = automatic (esn_cluster_auto_limit, from the measured crossover).
static int g_small_batch_limit = -2;
static int small_batch_setting() {
    if (g_small_batch_limit == -2) {
        const char *e = getenv("ESN_CLUSTER_MAX_B");
        g_small_batch_limit = e ? atoi(e) : -1;
        if (g_small_batch_limit < 0) g_small_batch_limit = -1;
    }
    return g_small_batch_limit;
}

extern "C" int esn_set_small_batch_limit(int max_frames) {
    const int old = small_batch_setting();
    g_small_batch_limit = max_frames < 0 ? -1 : max_frames;
    return old;
}

extern "C" int esn_pad_sizes(int N, int n_in, int n_out, int *N_pad, int *K_aug_pad) {
    if (N <= 0 || n_in <= 0 || n_out <= 0 || !N_pad || !K_aug_pad) return ESN_E_BADARG;
    *N_pad = (N + 127) / 128 * 128;
    *K_aug_pad = (N + n_in + n_out + 15) / 16 * 16;
    return 0;
}

extern "C" int esn_recurrence_run(const esn_recurrence_args *a, void *stream) {
    if (!a) return ESN_E_BADARG;
    if (a->B <= 0 || a->T <= 0 || a->N <= 0 || a->n_in <= 0 || a->n_out <= 0) return ESN_E_BADARG;
    if (a->n_out > ESN_MAX_OUT || a->n_in > ESN_MAX_IN) return ESN_E_BADARG;
    if (a->dtype != ESN_F32 && a->dtype != ESN_F64) return ESN_E_BADARG;
    int np_, kp_;
    esn_pad_sizes(a->N, a->n_in, a->n_out, &np_, &kp_);
    if (a->N_pad != np_ || a->K_aug_pad != kp_) return ESN_E_BADARG;
    if (!a->Wt_aug || !a->in || !a->in_scale || !a->in_shift || !a->t_scale || !a->t_shift ||
        !a->workspace)
        return ESN_E_BADARG;
    if (a->mode == ESN_MODE_HARVEST) {
        if (!a->ext_out || (a->feedback && !a->teacher)) return ESN_E_BADARG;
    } else if (a->mode == ESN_MODE_PREDICT) {
        if (!a->W_out || !a->y_out || a->n_groups <= 0 || a->transient < 0 || a->transient >= a->T)
            return ESN_E_BADARG;
    } else {
        return ESN_E_BADARG;
    }
    esn_simt::RecParams p;
    p.mode = a->mode; p.B = a->B; p.T = a->T; p.N = a->N; p.n_in = a->n_in; p.n_out = a->n_out;
    p.N_pad = a->N_pad; p.K_aug_pad = a->K_aug_pad; p.transient = a->transient;
    p.feedback = a->feedback; p.n_groups = a->n_groups; p.stage_wout = 0;
    p.noise_amp = a->noise_amp; p.seed = a->seed;
    p.Wt_aug = a->Wt_aug; p.in = a->in; p.in_scale = a->in_scale; p.in_shift = a->in_shift;
    p.teacher = a->teacher; p.t_scale = a->t_scale; p.t_shift = a->t_shift; p.W_out = a->W_out;
    p.group_ids = a->group_ids; p.x0 = a->x0; p.y0 = a->y0; p.noise = a->noise_uniforms;
    p.ext_out = a->ext_out; p.y_out = a->y_out; p.workspace = a->workspace;
    cudaStream_t st = (cudaStream_t)stream;
    int limit = small_batch_setting() >= 0 ? small_batch_setting() : esn_cluster_auto_limit(p, a->dtype);
    // the automatic limit is the crossover against the streaming SIMT kernel; fp64 runs of up to 512 neurons
    // continue on the fp64 tensor cores instead, which take 0.28 of that time (profiles/r2_fp64_harvest.txt)
    if (small_batch_setting() < 0 && a->dtype == ESN_F64 && a->N_pad <= 512 &&
        esn_dmma_enabled())
        limit = limit / 3;
    if (a->B <= limit) {
        const int rc = esn_cluster_launch(p, a->dtype, st);
        if (rc != ESN_E_UNSUPPORTED && rc != ESN_E_TOOLARGE) return rc;
    }
    if (a->dtype == ESN_F32) return esn_simt_launch_f32(p, st);
    {                                                // fp64, up to 1024 neurons: the fp64 tensor cores
        const int rc = esn_dmma_launch(p, st);
        if (rc != ESN_E_UNSUPPORTED) return rc;
    }
    return esn_simt_launch_f64(p, st);
}
