/*
 * esn_b200.h -- C ABI of the B200 (sm_100a) ESN symbol-detection engine.
 *
 * This is the drop-in boundary for the hot path of aoschu/esn-ofdm-mimo.  The
 * reference has no FFI of its own (pure Python); these entry
 * points are what a binding for its hot path would call, one per numerical
 * stage, each citing the reference lines it replaces (paths relative to the
 * reference root).  The Python drop-in modules (esn-ofdm-mimo_b200/libs/pyESN.py,
 * helper_mimo_esn_generic.py, HelpFunc.py) bind them through ctypes; see
 * INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless the name
 *    ends in `_host`; the caller owns all memory; nothing here allocates
 *    (except esn_ws_* helpers) or synchronises.
 *  - every launch function takes a `void *stream` (a cudaStream_t) and returns
 *    0 on success, a negative ESN_E_* for argument errors, or a positive
 *    cudaError_t.
 *  - matrices are row-major, frames are the slowest dimension.
 *  - `dtype`: ESN_F32 (fp32 SIMT FFMA path) or ESN_F64 (fp64 path).
 */
#ifndef ESN_B200_H
#define ESN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESN_F32 0
#define ESN_F64 1

#define ESN_E_BADARG   (-1)   /* null pointer / non-positive size / unsupported value */
#define ESN_E_TOOLARGE (-2)   /* shape exceeds what the kernel can stage on one SM */
#define ESN_E_NODEVICE (-3)   /* no sm_100 device */
#define ESN_E_UNSUPPORTED (-4) /* shape outside what this path supports (use the SIMT path) */

#define ESN_MAX_OUT 16        /* n_outputs <= 16 (2*N_t; the reference uses <= 8) */
#define ESN_MAX_IN  64        /* n_inputs  <= 64 (2*N_r; the reference uses <= 16) */

/* mode of esn_recurrence_run */
#define ESN_MODE_HARVEST 0    /* teacher-forced state harvesting of ESN.fit      */
#define ESN_MODE_PREDICT 1    /* free-running prediction of ESN.predict          */

int esn_version(void);
/* Fills name[0..n) with the device name and returns the SM count, or <0. */
int esn_device_info(char *name_host, int n, int *sm_major_host, int *sm_minor_host);
/* Host copy of the device state-noise stream: the uniform in [0,1) that the
 * recurrence kernels draw for (seed, frame, noise row, neuron). */
float esn_noise_uniform_host(unsigned long long seed, unsigned frame, unsigned row, unsigned neuron);

/* numpy's legacy generator on the device.  pyESN draws its state noise from `random_state_.rand(N_res)`, once
 * per time step (libs/pyESN.py:125): MT19937 through numpy's random_sample.  state_dev holds the generator's
 * state as get_state() returns it -- 624 key words followed by the position -- and is advanced in place by
 * 2 * count outputs; out receives the `count` uniforms rand() would have returned (dtype ESN_F64, or ESN_F32 =
 * the same doubles rounded); words_scratch: 2 * count 32-bit words, 8-byte aligned.  The host puts the advanced
 * state back with set_state(), so the caller's generator continues exactly where the reference's would. */
int esn_mt19937_uniforms(uint32_t *state_dev, long long count, int dtype, uint32_t *words_scratch, void *out,
                         void *stream);

/* ---------------------------------------------------------------------------
 * Reservoir recurrence.  Replaces the Python time loops and ESN._update:
 *   libs/pyESN.py:111-125 (_update), :179-182 (harvest loop of fit),
 *   :243-253 (free-running loop of predict), :127-152 (I/O scaling, folded in).
 *
 * Weights: Wt_aug[K_aug_pad][N_pad] (dtype), row k < N is column k of W,
 * rows N..N+n_in-1 are W_in^T, rows N+n_in..N+n_in+n_out-1 are W_feedb^T,
 * all other rows/columns zero.  N_pad is a multiple of 128, K_aug_pad a
 * multiple of 16 (esn_pad_sizes gives both).
 *
 * HARVEST: x_0 = 0; for n = 1..T-1: x_n = tanh(W x_{n-1} + W_in u_n + W_fb d_{n-1}) + noise_n,
 *   u_n = in[b,n,:]*in_scale+in_shift, d_n = teacher[b,n,:]*t_scale+t_shift.
 *   Writes the extended state E[b,n,:] = [x_n, u_n] for n = 0..T-1 into
 *   ext_out[B][T][N+n_in].   noise row consumed at step n is row n-1.
 * PREDICT: x_{-1} = x0[b] (or 0), y_{-1} = y0[b] (or 0); for n = 0..T-1:
 *   x_n = tanh(W x_{n-1} + W_in u_n + W_fb y_{n-1}) + noise_n,
 *   y_n = W_out[g(b)] [x_n; u_n];  y_out[b, n-transient, :] = (y_n - t_shift)/t_scale
 *   for n >= transient.  ext_out may be null (or receives E as above).
 * noise_n = noise_amp * (U - 0.5): U read from noise_uniforms[b][step][N]
 *   (dtype) when non-null, else drawn from a counter hash of (seed,b,n,neuron)
 *   when noise_amp != 0.  feedback = 0 drops the W_fb term (teacher_forcing=False).
 * ------------------------------------------------------------------------- */
typedef struct esn_recurrence_args {
    int32_t dtype;            /* ESN_F32 | ESN_F64 */
    int32_t mode;             /* ESN_MODE_* */
    int32_t B, T;             /* frames, time steps per frame */
    int32_t N, n_in, n_out;   /* reservoir size, inputs, outputs */
    int32_t N_pad, K_aug_pad; /* padded sizes of Wt_aug */
    int32_t transient;        /* PREDICT: first output row kept */
    int32_t feedback;         /* 1 = W_fb term on (teacher_forcing) */
    int32_t n_groups;         /* PREDICT: number of readouts in W_out */
    double  noise_amp;        /* ESN.noise */
    uint64_t seed;            /* device noise stream when noise_uniforms == null */
    const void *Wt_aug;       /* [K_aug_pad][N_pad] */
    const void *in;           /* [B][T][n_in] raw inputs */
    const void *in_scale;     /* [n_in] */
    const void *in_shift;     /* [n_in] */
    const void *teacher;      /* HARVEST: [B][T][n_out] raw teacher */
    const void *t_scale;      /* [n_out] */
    const void *t_shift;      /* [n_out] */
    const void *W_out;        /* PREDICT: [n_groups][n_out][N+n_in] */
    const int32_t *group_ids; /* PREDICT: [B] or null (all 0) */
    const void *x0;           /* PREDICT: [B][N] or null */
    const void *y0;           /* PREDICT: [B][n_out] (scaled domain) or null */
    const void *noise_uniforms; /* [B][T-1 | T][N] or null */
    void *ext_out;            /* [B][T][N+n_in] or null (required for HARVEST) */
    void *y_out;              /* PREDICT: [B][T-transient][n_out] */
    void *workspace;          /* [B][N] (dtype) scratch: the new state is parked here between steps */
} esn_recurrence_args;

int esn_pad_sizes(int N, int n_in, int n_out, int *N_pad_host, int *K_aug_pad_host);
int esn_recurrence_run(const esn_recurrence_args *args_host, void *stream);
/* esn_recurrence_run picks between two kernels.  Small batches with N <= 512 (the call shape of the
 * unmodified demos: one OFDM symbol per ESN.predict, one pilot per ESN.fit, libs/pyESN.py:154,218) run on one
 * thread-block cluster per 1-8 frames with [W | W_in | W_fb] resident in the cluster's shared memory and the
 * new state exchanged through DSMEM (~3 us per time step at N = 512); larger batches stream the weights from
 * L2 through the tiled SIMT kernel.  max_frames >= 0 fixes the largest batch the cluster kernel takes (0
 * disables it); a negative value restores the automatic choice (a multiple of the resident-cluster wave,
 * from the measured crossover; also the ESN_CLUSTER_MAX_B environment variable).  Returns the previous
 * setting (-1 = automatic). */
int esn_set_small_batch_limit(int max_frames);

/* ---------------------------------------------------------------------------
 * Tensor-core recurrence (tcgen05 + TMEM), free-running predict.  Same reference
 * lines as esn_recurrence_run in PREDICT mode (libs/pyESN.py:243-255).  fp16
 * hi/lo operand split (three MMAs per product, fp32 accumulation in TMEM), 64
 * frames per CTA; reservoirs padded to 256 or 512 neurons run on CTA pairs
 * (cta_group::2, 128 frames per pair, state halves exchanged through DSMEM).
 * Frames share a readout in aligned runs: group_ids must be uniform over every
 * aligned 64 frames when n_out <= 8 (the 16 readout rows of the MMA then carry one
 * readout per CTA of the pair), over every aligned 128 frames otherwise.
 * N <= 512, n_in <= 24, n_out <= 16.
 *
 * esn_tc_prepare_weights builds the UMMA-ready (pre-swizzled, fp16 hi/lo) image
 * of [W | W_in | 0 | W_feedb] once per reservoir -- it is shared by all CTAs and
 * all readouts.  esn_tc_prepare_readout builds the small per-readout image of
 * W_out (and yscale[g]).  Power-of-two pre-scales keep the fp16 halves in range:
 *   su_exp: inputs (after in_scale/in_shift) are multiplied by 2^su_exp; choose
 *           it so that max|u| 2^su_exp is about 2^9 (su_exp >= 1);
 *   sy_exp: fed-back outputs (scaled teacher domain) are multiplied by 2^sy_exp;
 *           choose it so that max|y| 2^sy_exp is about 2^6.
 * ------------------------------------------------------------------------- */
typedef struct esn_tc_predict_args {
    int32_t B, T;
    int32_t N, n_in, n_out;
    int32_t transient;
    int32_t feedback;           /* 1 = teacher_forcing (W_feedb y fed back) */
    int32_t su_exp, sy_exp;
    int32_t n_groups;
    double  noise_amp;
    uint64_t seed;
    const void *weights;        /* from esn_tc_prepare_weights (same su_exp, sy_exp) */
    const void *readouts;       /* [n_groups][esn_tc_readout_bytes], from esn_tc_prepare_readout */
    const float *yscale;        /* [n_groups], from esn_tc_prepare_readout */
    const float *in;            /* [B][T][n_in] raw inputs (fp32) */
    const float *in_scale, *in_shift;   /* [n_in] */
    const float *t_scale, *t_shift;     /* [n_out] */
    const int32_t *group_ids;   /* [B] or null */
    const float *x0;            /* [B][N] initial state or null (zeros) */
    const float *y0;            /* [B][n_out] initial output (scaled domain) or null */
    const float *noise_uniforms;/* [B][T][N] or null (device counter stream) */
    float *ext_out;             /* [B][T][N+n_in] or null */
    float *y_out;               /* [B][T-transient][n_out] */
    void *timeline;             /* profiling aid: [T+1][8] int64 SM-clock stamps of CTA 0, or null */
    int32_t reserved;           /* must be 0 */
    const float *teacher;       /* harvest mode: [B][T][n_out] raw teachers; the recurrence is
                                   teacher-forced as in ESN.fit (libs/pyESN.py:179-182), ext_out [B][T][N+n_in] is
                                   required, readouts / yscale / y_out are ignored.  null = free-running predict */
} esn_tc_predict_args;

int esn_tc_supported(int N, int n_in, int n_out);
long long esn_tc_weight_bytes(int N, int n_in);
long long esn_tc_readout_bytes(int N, int n_in);
int esn_tc_prepare_weights(const double *W, const double *W_in, const double *W_fb, int N, int n_in,
                           int n_out, int su_exp, int sy_exp, int feedback, void *image, void *stream);
int esn_tc_prepare_readout(const double *W_out, int N, int n_in, int n_out, int n_groups, int su_exp,
                           void *image, float *yscale, void *stream);
int esn_tc_predict(const esn_tc_predict_args *args_host, void *stream);

/* Truncation-bias compensation of the tensor-core paths (esn_tc_predict, esn_tcs_run).  The tensor core adds each
 * K = 16 product block to its fp32 accumulator with truncation toward zero, so a chain of n MMAs comes out short
 * by about n x k0 relative; the recurrence amplifies that systematic shrink of W x (libs/pyESN.py:117-119) by
 * ~1 / (1 - spectral radius).  The epilogues scale the accumulators by 1 + n k0.  k0 >= 0 sets the per-MMA loss
 * (0 = no compensation), a negative value restores the calibrated default (profiles/r2_tc_acc_bias.txt).
 * Returns the previous value.  esn_tc_acc_k0 reads it. */
double esn_tc_set_acc_k0(double k0);
double esn_tc_acc_k0(void);

/* ---------------------------------------------------------------------------
 * Tensor-core recurrence with the state streamed through L2 ("tcs"): the same
 * reference lines and the same fp16 hi/lo arithmetic as esn_tc_predict
 * (libs/pyESN.py:179-182, 243-255), for what the resident kernel cannot take:
 *   - reservoirs of 513..4096 neurons (the 4x8 fast demo's 600 neurons,
 *     system_model_2/Demo_MIMO_4x8_ChannelRank_TrainSNR_LDPC_fast.py:142; the
 *     sweep's 1024 / 2048, BASELINE.json configs[3]);
 *   - ANY frame -> readout map (group_ids unrestricted): the demos train a new
 *     readout every L = 19 symbols (OFDM_MIMO_2-2_NBF_LDPC.py:151-153, 270), so
 *     18 consecutive data frames share a W_out.
 * The state of a CTA's 64 frames lives in `workspace` (esn_tcs_workspace_bytes)
 * as UMMA-ready fp16 hi/lo tiles, double-buffered over time steps; the readout
 * W_out[g(b)] [x; u] runs on the CUDA cores from fp32 tables built by
 * esn_tcs_prepare_readout (wo_x: esn_tcs_readout_floats() floats per readout,
 * wo_u: *wo_u_floats_host per readout).  The weight image is the one
 * esn_tc_prepare_weights builds (it accepts N <= 4096).  n_in <= 24, n_out <= 16.
 * accumulators (0 = automatic, 2 or 4): TMEM accumulators per 256-neuron pass -- the
 * correction products and the main products of every (accumulators - 1)-th chunk are
 * summed separately and added in fp32 RN by the epilogue, which keeps the tensor core's
 * truncating accumulate chains short (2: two TMEM buffers, 4: one);
 * ring_a / ring_b (0 = automatic): depths of the state / weight rings.
 * ------------------------------------------------------------------------- */
typedef struct esn_tcs_args {
    int32_t B, T;
    int32_t N, n_in, n_out;
    int32_t transient;
    int32_t feedback;
    int32_t su_exp, sy_exp;     /* as esn_tc_predict_args */
    int32_t n_groups;
    int32_t accumulators, ring_a, ring_b;
    double  noise_amp;
    uint64_t seed;
    const void *weights;        /* from esn_tc_prepare_weights (same su_exp, sy_exp) */
    const float *wo_x;          /* [n_groups][esn_tcs_readout_floats] */
    const float *wo_u;          /* [n_groups][wo_u floats] */
    const float *in;            /* [B][T][n_in] raw inputs (fp32) */
    const float *in_scale, *in_shift;   /* [n_in] */
    const float *t_scale, *t_shift;     /* [n_out] */
    const int32_t *group_ids;   /* [B] or null; any values in [0, n_groups) (clamped) */
    const float *x0;            /* [B][N] or null */
    const float *y0;            /* [B][n_out] (scaled domain) or null */
    const float *noise_uniforms;/* [B][steps][N] or null (device counter stream) */
    float *ext_out;             /* [B][T][N+n_in] or null (required in harvest mode) */
    float *y_out;               /* [B][T-transient][n_out] (predict mode) */
    const float *teacher;       /* harvest mode: [B][T][n_out] raw teachers, else null */
    void *workspace;            /* esn_tcs_workspace_bytes(B, N) bytes */
    void *timeline;             /* profiling aid: [steps][16] int64 SM-clock stamps of CTA 0, or null */
} esn_tcs_args;

int esn_tcs_supported(int N, int n_in, int n_out);
long long esn_tcs_workspace_bytes(int B, int N);
long long esn_tcs_readout_floats(int N, int n_out, long long *wo_u_floats_host);
int esn_tcs_prepare_readout(const double *W_out, int N, int n_in, int n_out, int n_groups, float *wo_x,
                            float *wo_u, void *stream);
int esn_tcs_run(const esn_tcs_args *args_host, void *stream);

/* The same recurrence with the state RESIDENT in shared memory ("tcr", reservoirs of up to 512 neurons, n_in <= 16,
 * n_out <= 8): the machine of esn_tc_predict (CTA pair, frames on M, weights streamed through a ring) with the
 * readout on the CUDA cores in fp32 round-to-nearest and split accumulators.  Takes the argument block of
 * esn_tcs_run; workspace, accumulators, ring_a and ring_b are ignored; timeline is [T + 33][8] int64.
 * group_ids must be uniform over aligned runs of 64 frames (one readout per CTA; the readout of a run's FIRST
 * frame is used). */
int esn_tcr_supported(int N, int n_in, int n_out);
int esn_tcr_run(const esn_tcs_args *args_host, void *stream);

/* ---------------------------------------------------------------------------
 * Readout training.  Replaces np.linalg.pinv + dot of ESN.fit
 * (libs/pyESN.py:189-192) by fp64 normal equations with lambda = 0:
 *   rows m = T - transient, cols p = N + n_in, E = ext[b, transient:, :]
 *   m >= p (primal): G = E^T E [p x p], R = E^T D [p x n_out], W_out^T = G^-1 R
 *   m <  p (dual)  : G = E E^T [m x m], A = G^-1 D,            W_out^T = E^T A
 * which equals the pinv (min-norm / least-squares) solution when E has full
 * rank.  All accumulation in fp64; `ext` (the extended states written by
 * esn_recurrence_run) and `teacher` may be fp32 or fp64 (ext_dtype/teacher_dtype).  D = teacher[b, transient:, :]*t_scale + t_shift.
 *
 * esn_gram_f64: builds G (lower triangle + mirrored) and the right-hand side
 *   rhs[b] = R (primal, [p][n_out]) or D (dual, [m][n_out]).
 *   If `accumulate` != 0, G and rhs are ADDED into (shared-readout training over
 *   many frames, primal only); `shared` != 0 sums all B frames into problem 0.
 * esn_cholesky_solve_f64: in-place batched Cholesky G = L L^T and solve for
 *   n_rhs right-hand sides; info[b] = 0 or the 1-based index of the first
 *   non-positive pivot (as LAPACK potrf).
 * esn_cholesky_solve_piv_f64: the same, and pivots[b] = {min_j d_jj, max_j d_jj} (d_jj = l_jj^2, the
 *   pivots of the factorisation): max / min is a lower bound of cond(G) = cond(E)^2.  The reference's
 *   pinv (libs/pyESN.py:191-192) is an SVD solve and survives cond(E) ~ 1e9 where lambda = 0 normal
 *   equations lose all digits without failing; the host uses the ratio to route such problems to a
 *   stable solve.  `pivots` may be null.
 * esn_readout_from_dual_f64: W_out[b] = (E^T A)^T, [n_out][p].
 * esn_transpose_rhs_f64: primal: W_out[b] = rhs^T.
 * ------------------------------------------------------------------------- */
int esn_gram_f64(const void *ext, int ext_dtype, const void *teacher, int teacher_dtype,
                 const double *t_scale, const double *t_shift,
                 int B, int T, int p, int n_out, int transient,
                 int dual, int shared, int accumulate,
                 double *G, double *rhs, void *stream);
int esn_cholesky_solve_f64(double *G, double *rhs, int batch, int n, int n_rhs,
                           int32_t *info, void *stream);
int esn_cholesky_solve_piv_f64(double *G, double *rhs, int batch, int n, int n_rhs,
                               int32_t *info, double *pivots, void *stream);
int esn_readout_from_dual_f64(const void *ext, int ext_dtype, const double *A, int B, int T,
                              int p, int n_out, int transient, double *W_out, void *stream);
int esn_transpose_rhs_f64(const double *rhs, int B, int p, int n_out, double *W_out,
                          void *stream);

/* Train-set prediction of ESN.fit (libs/pyESN.py:212-213):
 * pred[b,n,:] = (W_out[g(b)] E[b,n,:] - t_shift)/t_scale for all T rows. */
int esn_apply_readout(int dtype, const void *ext, const void *W_out, const int32_t *group_ids,
                      const void *t_scale, const void *t_shift,
                      int B, int T, int p, int n_out, void *pred, void *stream);

/* ---------------------------------------------------------------------------
 * OFDM side (north-star subsystem 3).
 * ofdm_unpack_fft_demap: ESN time-domain output -> per-Tx complex sequence ->
 *   (1/N) FFT_N / sqrt(Pi) -> hard QAM decision -> bit errors.  Replaces
 *   system_model_2/OFDM_MIMO_2-2_NBF_LDPC.py:56-64 (reconstruct), :435-438
 *   (FFT), :103-111 + :36-38 (hard_bits_from_syms / bits_to_grayvec),
 *   :469-474 (error count).
 *   y [B][rows][2*N_t] (dtype; first N rows used), Pi[B] or Pi[1] (pi_stride 0/1),
 *   X_hat [B][N][N_t] complex (2 x dtype) or null, idx [B][N][N_t] u8 or null,
 *   tx_idx [B][N][N_t] u8 or null (transmitted symbol indices) -> bit errors
 *   added to err_count[0] (u64) and near-boundary symbols (|distance to a
 *   slicer boundary| < boundary_eps) to err_count[1].
 * ofdm_rx_fft: Y = (1/N) FFT(y_CP[CP:]) (:428).  y_cp [B][N+CP][N_r] complex.
 * ofdm_equalize: per-subcarrier solve(H^H H + reg I, H^H Y)/power_scale
 *   (:41-53, :453-460).  H [Bh][N][N_r][N_t] complex with h_index[b] (or null
 *   = b) selecting the estimate of the frame's coherence block.
 * ofdm_chanest: pilot LS on the comb tx::N_t, linear inter/extrapolation,
 *   IFFT, truncate to `taps`, (mmse_scaler*R_h^-1 + I)^-1, FFT (:316-334).
 * ofdm_demap_count: slicer + error count on an arbitrary X_hat (for ZF/MMSE).
 * N must be a power of two <= 4096; qam_bits in {2,4,6}.
 * ------------------------------------------------------------------------- */
int ofdm_unpack_fft_demap(int dtype, const void *y, int B, int rows, int N, int N_t,
                          const void *Pi, int pi_stride, int qam_bits,
                          void *X_hat, uint8_t *idx, const uint8_t *tx_idx,
                          double boundary_eps, unsigned long long *err_count,
                          void *stream);
int ofdm_rx_fft(int dtype, const void *y_cp, int B, int N, int cp, int N_r,
                void *Y, void *stream);
int ofdm_equalize(int dtype, const void *Y, const void *H, const int32_t *h_index,
                  int B, int N, int N_r, int N_t, const void *reg, int reg_stride,
                  const void *power_scale, int ps_stride, void *X_hat, void *stream);
int ofdm_chanest(int dtype, const void *Y_LS, const void *X_LS, int B, int N, int N_r,
                 int N_t, const void *Pi, const void *isi_magnitude, int taps,
                 double No, void *H_LS, void *H_MMSE, void *stream);
/* Workload generation: symbol indices -> QAM -> N*IFFT -> CP -> *sqrt(Pi)
 * -> soft PA clip x/sqrt(1+(|x|/A)^2) -> FIR channel per link -> + noise_std * CN(0,2) -> y_cp
 * [B][N+CP][N_r] complex and/or the real ESN input rows esn_in [B][N+CP+delay][2 N_r].
 * Replaces system_model_2/OFDM_MIMO_2-2_NBF_LDPC.py:402-426 (Tx chain, lfilter, AWGN) and
 * :430-433 (ESN input packing).  taps [n_chan][N_r][N_t][ntaps] complex, chan_index[b] or null
 * (= b); a tx_idx of 255 is an empty subcarrier (comb pilots, :287-289);
 * noise [B][N+CP][N_r] complex standard normals or null (device counter stream `seed`);
 * x_cp [B][N+CP][N_t] complex receives the unclipped Tx samples (the ESN teacher) or null. */
int ofdm_synth_frames(int dtype, const uint8_t *tx_idx, const void *taps, const int32_t *chan_index,
                      const void *Pi, const void *A_clip, const void *noise, double noise_std,
                      unsigned long long seed, int B, int N, int cp, int N_t, int N_r, int ntaps,
                      int qam_bits, int delay, void *x_cp, void *y_cp, void *esn_in, void *stream);
int ofdm_demap_count(int dtype, const void *X_hat, int B, int N, int N_t, int qam_bits,
                     uint8_t *idx, const uint8_t *tx_idx, double boundary_eps,
                     unsigned long long *err_count, void *stream);
/* Soft outputs, replacing system_model_2/Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:
 * ofdm_soft_demap: per frame sigma2 = mean over Tx of (mean |X_hat - nearest point|^2 + 1e-12)
 *   (est_sigma2_from_decision :84-88, :459), then max-log LLRs (d1 - d0) / max(sigma2, 1e-12) per bit,
 *   bit b of the point index LSB first (qam_llrs_maxlog :66-82, bits_to_grayvec :30-32); positive =
 *   bit 0.  With cal_a / cal_b [qam_bits] (fp64, device) the decoder-side map clip(-(a_b llr + b_b),
 *   +-clip) is applied (:489-492).  X_hat [B][N][N_t] complex, sigma2 [B] or null, llr
 *   [B][N][qam_bits][N_t] (the reference's (N, m, N_t) per frame) or null.
 * ofdm_llr_calibrate: fit_logreg_1d (:108-119) per bit position over every symbol of the B frames:
 *   maxiter steps of full-batch gradient descent (rate lr, ridge l2 on a) from a = 1, b = 0; y = bit of
 *   tx_idx [B][N][N_t].  ab [qam_bits][2] fp64 (device). */
int ofdm_soft_demap(int dtype, const void *X_hat, int B, int N, int N_t, int qam_bits,
                    const double *cal_a, const double *cal_b, double clip, void *sigma2, void *llr,
                    void *stream);
int ofdm_llr_calibrate(int dtype, const void *llr, const uint8_t *tx_idx, int B, int N, int N_t,
                       int qam_bits, int maxiter, double lr, double l2, double *ab, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ESN_B200_H */
