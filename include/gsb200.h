/*
 * gsb200.h -- C ABI of the B200-native randomisation-method summator.
 *
 * This is the drop-in boundary for the ONE hot path of GSTools: the native
 * `summate` / `summate_incompr` functions that gstools imports from gstools-cython /
 * gstools_core and dispatches through `_summate` / `_summate_incompr`.
 * Every entry point cites the reference interface it replaces (paths relative to
 * the reference repository root).
 *
 * Conventions
 *   - All arrays are fp64, row-major.  `cov_samples` is (dim, n_modes), `z_1`/`z_2`
 *     are (n_modes,), `pos` is (dim, n_pts) with row stride `pos_ld` (elements),
 *     scalar output is (n_pts,), vector output is (dim, n_pts) with row stride `out_ld`.
 *   - `mem` says where EVERY pointer of the call lives: GSB_MEM_HOST (the library
 *     stages host<->device copies itself) or GSB_MEM_DEVICE (pointers are device
 *     pointers on `device`, nothing is copied; work is enqueued on `stream` and the
 *     call returns without synchronising).
 *   - Inputs are never modified.  The caller owns all buffers.
 *   - Return value: 0 on success, non-zero on error; gsb_last_error() then returns a
 *     thread-local message.  No exception crosses this boundary.
 *   - The sqrt(var/N) scale and the nugget are NOT applied here; as in the reference
 *     they belong to the caller (src/gstools/field/generator.py:269-270, 561-567).
 *   - There is no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef GSB200_H
#define GSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_MEM_HOST 0
#define GSB_MEM_DEVICE 1

#define GSB_OK 0
#define GSB_ERR_ARGUMENT 1
#define GSB_ERR_CUDA 2
#define GSB_ERR_NO_DEVICE 3

#define GSB_MAX_DIM 8
#define GSB_EPI_MAX_ADD 4
#define GSB_EPI_MAX_COMP 3

/*
 * Fused caller epilogue.  The reference applies, on the host and
 * one numpy pass each, to the array the native function returns:
 *     RandMeth.__call__        sqrt(var/N) * summed + nugget          generator.py:269-270
 *     IncomprRandMeth.__call__ mean_u*e1 + mean_u*sqrt(var/N)*summed + nugget   generator.py:561-567
 *     apply_mean_norm_trend    field += mean; (identity normalizer); field += trend
 *                                                                     normalizer/tools.py:99-103
 * With a gsb_epilogue the kernels store
 *     v = scale * sum;  v = v + add[0][c];  v = v + add[1][c]; ...  (n_add terms, c = component)
 * with separately rounded operations in exactly this order, i.e. the bits the numpy passes give
 * for constant nugget / mean / trend.  NULL means "raw sums" (the reference's native function).
 */
typedef struct gsb_epilogue {
    double scale;
    int32_t n_add;      /* 0..GSB_EPI_MAX_ADD */
    int32_t reserved;
    double add[GSB_EPI_MAX_ADD][GSB_EPI_MAX_COMP];
} gsb_epilogue;

/* Library version (major*10000 + minor*100 + patch). */
int gsb_version(void);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char *gsb_last_error(void);

/* Number of visible CUDA devices (0 when there is none; that is not an error). */
int gsb_device_count(int *count);

/*
 * gsb_summate -- replaces gstools_cython.field.summate / gstools_core.summate
 *   imported at src/gstools/field/generator.py:22,32; called at generator.py:48
 *   as summate_fct(cov_samples, z_1, z_2, pos, num_threads).
 *   out[i] = sum_j z_1[j] cos(k_j . x_i) + z_2[j] sin(k_j . x_i)     (generator.py:193-199)
 * Unstructured ("direct") kernel; any dim in 1..GSB_MAX_DIM.
 */
int gsb_summate(const double *cov_samples, const double *z_1, const double *z_2,
                const double *pos, int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts,
                double *out, int mem, int device, void *stream);

/*
 * gsb_summate_incompr -- replaces gstools_cython.field.summate_incompr /
 *   gstools_core.summate_incompr (generator.py:24,34; called at generator.py:64).
 *   out[t,i] = sum_j p_t(k_j) (z_1[j] cos(k_j.x_i) + z_2[j] sin(k_j.x_i)),
 *   p_t(k) = delta_{t0} - k_t k_0 / |k|^2                             (generator.py:479-495)
 */
int gsb_summate_incompr(const double *cov_samples, const double *z_1, const double *z_2,
                        const double *pos, int64_t pos_ld, int dim, int64_t n_modes,
                        int64_t n_pts, double *out, int64_t out_ld, int mem, int device,
                        void *stream);

/*
 * gsb_summate_structured[_incompr] -- the same sums on a structured (rectilinear) mesh
 * WITHOUT the flat (dim, n) position array.  The reference expands the mesh on the host
 * (src/gstools/field/base.py:289-290 -> tools/geometric.py:340-356 generate_grid) and
 * isometrises it (base.py:297 -> covmodel/base.py:572-582, pos_iso = matrix @ pos) before
 * calling summate; this entry takes the axes and the (dim x dim) isometrisation matrix
 * instead (matrix == NULL means identity) and evaluates
 *     out[i_0,...,i_{d-1}] = summate(cov_samples, z_1, z_2, matrix @ (axes_0[i_0],...))
 * in C order (last axis fastest), i.e. exactly the array the reference reshapes at
 * src/gstools/field/srf.py:156.
 *   axes      : all axis coordinates concatenated, sum(axis_len) doubles
 *   axis_len  : dim entries
 *   n_batch   : number of independent mode sets evaluated on the same mesh
 *               (ensembles: examples/06_conditioned_fields/01_2D_condition_ensemble.py:32-35);
 *               cov_samples is (n_batch, dim, n_modes), z_1/z_2 are (n_batch, n_modes),
 *               out is (n_batch, n) resp. (n_batch, dim, n).  n_batch = 1 for a single field.
 */
int gsb_summate_structured(const double *cov_samples, const double *z_1, const double *z_2,
                           const double *axes, const int64_t *axis_len, const double *matrix,
                           int dim, int64_t n_modes, int64_t n_batch, double *out, int mem,
                           int device, void *stream);

int gsb_summate_incompr_structured(const double *cov_samples, const double *z_1,
                                   const double *z_2, const double *axes,
                                   const int64_t *axis_len, const double *matrix, int dim,
                                   int64_t n_modes, int64_t n_batch, double *out, int mem,
                                   int device, void *stream);

/*
 * The same four entry points with the caller epilogue fused into the kernels' stores
 * (`epi` == NULL: identical to the functions above).
 */
int gsb_summate_ex(const double *cov_samples, const double *z_1, const double *z_2,
                   const double *pos, int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts,
                   double *out, const gsb_epilogue *epi, int mem, int device, void *stream);

int gsb_summate_incompr_ex(const double *cov_samples, const double *z_1, const double *z_2,
                           const double *pos, int64_t pos_ld, int dim, int64_t n_modes,
                           int64_t n_pts, double *out, int64_t out_ld, const gsb_epilogue *epi,
                           int mem, int device, void *stream);

int gsb_summate_structured_ex(const double *cov_samples, const double *z_1, const double *z_2,
                              const double *axes, const int64_t *axis_len, const double *matrix,
                              int dim, int64_t n_modes, int64_t n_batch, double *out,
                              const gsb_epilogue *epi, int mem, int device, void *stream);

int gsb_summate_incompr_structured_ex(const double *cov_samples, const double *z_1,
                                      const double *z_2, const double *axes,
                                      const int64_t *axis_len, const double *matrix, int dim,
                                      int64_t n_modes, int64_t n_batch, double *out,
                                      const gsb_epilogue *epi, int mem, int device, void *stream);

/*
 * Per-point affine epilogue for CONDITIONED fields (second half).
 * CondSRF.__call__ combines the random field with the kriging results on the host, one numpy pass
 * each over the whole mesh (src/gstools/field/cond_srf.py:133-150, get_scaling :152-178):
 *     var_scale = sqrt(krige_var / var)                                   cond_srf.py:176
 *     field     = rawkrige + var_scale * rawfield + nugget                cond_srf.py:146
 * followed by post_field's constant mean / trend (normalizer/tools.py:99-103).  With a
 * gsb_point_epilogue the kernels store, after the terms of gsb_epilogue (which yield `rawfield`),
 *     v = gain[i] * v;  v = offset[i] + v;  v = v + add[0];  v = v + add[1]; ...
 * with separately rounded operations in this order (gain = var_scale, offset = rawkrige; i = index
 * of the point in its field, C order for meshes), i.e. the bits of the numpy passes.  `gain` and
 * `offset` are ALWAYS device pointers on `device` (they are the device-resident results of one
 * kriging evaluation, reused by every realisation of an ensemble), whatever `mem` says about the
 * other arguments; either may be NULL (gain 1 / offset 0: the step is skipped).  Scalar fields only
 * (CondSRF.valid_value_types, cond_srf.py:56).  With n_batch > 1 every field of the batch uses the same arrays.
 */
typedef struct gsb_point_epilogue {
    const double *gain;     /* (n_pts,) device memory, or NULL */
    const double *offset;   /* (n_pts,) device memory, or NULL */
    int32_t n_add;          /* 0..GSB_EPI_MAX_ADD */
    int32_t reserved;
    double add[GSB_EPI_MAX_ADD];
} gsb_point_epilogue;

int gsb_summate_pp(const double *cov_samples, const double *z_1, const double *z_2,
                   const double *pos, int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts,
                   double *out, const gsb_epilogue *epi, const gsb_point_epilogue *pepi, int mem,
                   int device, void *stream);

int gsb_summate_structured_pp(const double *cov_samples, const double *z_1, const double *z_2,
                              const double *axes, const int64_t *axis_len, const double *matrix,
                              int dim, int64_t n_modes, int64_t n_batch, double *out,
                              const gsb_epilogue *epi, const gsb_point_epilogue *pepi, int mem,
                              int device, void *stream);

/*
 * gsb_cond_scaling -- CondSRF.get_scaling for a model without nugget (cond_srf.py:175-177) plus the
 * variance clamp of Krige.__call__ (krige/base.py:296-298), on the device:
 *     krige_var[i] = max(sill - error[i], 0);   gain[i] = sqrt(krige_var[i] / var)
 * `error` is the second output of gsb_krige_evaluate / gsb_calc_field_krige_and_variance.  All pointers
 * are device pointers; krige_var may be NULL or alias `error`.  IEEE subtraction, division and square
 * root: the same bits as numpy.
 */
int gsb_cond_scaling(const double *error, int64_t n, double sill, double var, double *krige_var,
                     double *gain, int device, void *stream);

/*
 * gsb_summate_fourier[_structured] -- replaces gstools_cython.field.summate_fourier /
 *   gstools_core.summate_fourier (imported generator.py:23,33; called generator.py:67-75 from the
 *   Fourier generator, generator.py:685-692):
 *   out[i] = sum_j spectrum_factor[j] (z_1[j] cos(k_j . x_i) + z_2[j] sin(k_j . x_i)),
 *   `modes` is (dim, n_modes).  Same kernels with the factor folded into the mode weights on the
 *   device.  
 */
int gsb_summate_fourier(const double *spectrum_factor, const double *modes, const double *z_1,
                        const double *z_2, const double *pos, int64_t pos_ld, int dim,
                        int64_t n_modes, int64_t n_pts, double *out, int mem, int device,
                        void *stream);

int gsb_summate_fourier_structured(const double *spectrum_factor, const double *modes,
                                   const double *z_1, const double *z_2, const double *axes,
                                   const int64_t *axis_len, const double *matrix, int dim,
                                   int64_t n_modes, double *out, int mem, int device,
                                   void *stream);

/*
 * gsb_calc_field_krige_and_variance / gsb_calc_field_krige -- replace
 *   gstools_cython.krige.calc_field_krige_and_variance / calc_field_krige (and the gstools_core
 *   twins), imported at src/gstools/krige/base.py:16-19, 30-33, dispatched at base.py:42-61 and
 *   called from Krige._summate (base.py:307-317) as fct(krig_mat, krig_vecs, cond, num_threads):
 *     field[k] = sum_i cond[i] * (krig_mat @ krig_vecs)[i,k]
 *     error[k] = sum_i krig_vecs[i,k] * (krig_mat @ krig_vecs)[i,k]
 *   krig_mat  (krige_size, krige_size) row-major     the inverted kriging matrix, base.py:319-357
 *   krig_vecs (krige_size, n_pts), row stride vecs_ld right-hand sides of one chunk, base.py:359-388
 *   cond      (krige_size,)                           base.py:562-565
 *   field, error (n_pts,)
 *   Host buffers of any size are streamed through the device
 * in column chunks (option "krige_host_chunk_mb").
 */
int gsb_calc_field_krige_and_variance(const double *krig_mat, const double *krig_vecs,
                                      int64_t vecs_ld, const double *cond, int64_t krige_size,
                                      int64_t n_pts, double *field, double *error, int mem,
                                      int device, void *stream);

int gsb_calc_field_krige(const double *krig_mat, const double *krig_vecs, int64_t vecs_ld,
                         const double *cond, int64_t krige_size, int64_t n_pts, double *field,
                         int mem, int device, void *stream);

/*
 * gsb_krige_evaluate[_structured] -- the whole evaluation loop of Krige.__call__
 * (src/gstools/krige/base.py:278-294) on the device: the right-hand sides of
 * Krige._get_krige_vecs (base.py:359-388) are generated there instead of being built on the host
 * (cdist + covariance: K x n doubles, 16.8 GB for 1000 conditioning points on a 128^3 mesh) and are
 * contracted straight away as above.
 *   model      covariance model: var * cor(r / len_rescaled), or `sill` at r ~ 0 when `exact`
 *              (CovModel.covariance / cov_nugget, covmodel/tools.py:65-76, covmodel/base.py:313-320);
 *              `type` selects cor(h) among the closed forms of covmodel/models.py
 *   cond_pos   (dim, cond_no) isometrised conditioning positions (Krige._krige_pos, base.py:584)
 *   pos        (dim, n_pts) isometrised evaluation positions, row stride pos_ld    [flat variant]
 *   axes, axis_len, matrix   structured mesh as in gsb_summate_structured          [structured variant]
 *   unbiased   1: row cond_no of the right-hand side is all ones (base.py:377-378)
 *   tail_rows  (krige_size - cond_no - unbiased, n_pts), row stride tail_ld: functional and external
 *              drift rows evaluated by the caller (base.py:379-387); NULL when there are none
 *   error      NULL: field only (return_var=False)
 */
#define GSB_COV_GAUSSIAN 1
#define GSB_COV_EXPONENTIAL 2
#define GSB_COV_STABLE 3      /* param = alpha */
#define GSB_COV_RATIONAL 4    /* param = alpha */
#define GSB_COV_CUBIC 5
#define GSB_COV_LINEAR 6
#define GSB_COV_CIRCULAR 7
#define GSB_COV_SPHERICAL 8

typedef struct gsb_cov_model {
    int32_t type;
    int32_t exact;
    double var;
    double len_rescaled;
    double sill;
    double param;
} gsb_cov_model;

int gsb_krige_evaluate(const gsb_cov_model *model, const double *krig_mat, const double *cond,
                       int64_t krige_size, const double *cond_pos, int64_t cond_no, int dim,
                       const double *pos, int64_t pos_ld, int64_t n_pts, int unbiased,
                       const double *tail_rows, int64_t tail_ld, double *field, double *error,
                       int mem, int device, void *stream);

int gsb_krige_evaluate_structured(const gsb_cov_model *model, const double *krig_mat,
                                  const double *cond, int64_t krige_size, const double *cond_pos,
                                  int64_t cond_no, int dim, const double *axes,
                                  const int64_t *axis_len, const double *matrix, int unbiased,
                                  const double *tail_rows, int64_t tail_ld, double *field,
                                  double *error, int mem, int device, void *stream);

/*
 * gsb_sample_radii_mcmc -- host-side, stream-compatible restatement of RNG.sample_ln_pdf
 * (src/gstools/random/rng.py:38-104): the emcee ensemble sampler
 * (stretch move, a = 2) that draws the mode radii of RandMeth for models without an inverse CDF
 * (generator.py:381-385).  Runs `burn_in` steps from the numpy legacy MT19937 state
 * (mt_key_burn[624], mt_pos_burn), then `n_steps` production steps from (mt_key_main, mt_pos_main) --
 * the reference hands each run_mcmc call the state of a fresh RandomState (rng.py:84-99, 193-203) --
 * and writes the positions of all walkers after every production step to chain[n_steps][nwalkers]
 * (= get_chain(flat=True)).
 *   pdf_kind   GSB_PDF_EXPONENTIAL / GSB_PDF_MATERN / GSB_PDF_GAUSSIAN: CovModel.ln_spectral_rad_pdf of that model
 *              (covmodel/base.py:553-560, covmodel/tools.py:374-406, models.py:217-224, 434-449)
 *   init       nwalkers initial positions (rng.py:78-80)
 * No device is touched.
 */
#define GSB_PDF_EXPONENTIAL 1
#define GSB_PDF_MATERN 2
#define GSB_PDF_GAUSSIAN 3
int gsb_sample_radii_mcmc(int pdf_kind, int dim, double len_rescaled, double nu,
                          const uint32_t *mt_key_burn, int mt_pos_burn,
                          const uint32_t *mt_key_main, int mt_pos_main, const double *init,
                          int nwalkers, int burn_in, int n_steps, double *chain);

/*
 * The same chain for ANY model: the log-pdf is the caller's -- `ln_pdf(r, n, out, user)` writes the log of the radial
 * spectral density of n radii (the reference hands emcee `model.ln_spectral_rad_pdf` with vectorize=True: one
 * evaluation per half ensemble, src/gstools/random/rng.py:88, covmodel/base.py:557-560) and returns 0, or non-zero to
 * abort.  The stretch move, the red/blue split and the consumption of the MT19937 stream are native; models without
 * a native closed form (Integral, HyperSpherical, JBessel, the TPL family, and every model whose spectral density is
 * a numerical Hankel transform) keep their own Python density and still lose emcee's per-step Python overhead.
 */
typedef int (*gsb_ln_pdf_fn)(const double *r, int n, double *ln_pdf, void *user);
int gsb_sample_radii_mcmc_cb(gsb_ln_pdf_fn ln_pdf, void *user, const uint32_t *mt_key_burn, int mt_pos_burn,
                             const uint32_t *mt_key_main, int mt_pos_main, const double *init, int nwalkers,
                             int burn_in, int n_steps, double *chain);

/*
 * gsb_sample_modes_batch -- the random streams of RandMeth.reset_seed (src/gstools/field/generator.py:346-387) for
 * MANY seeds at once, one seed per task on `n_threads` host threads: an ensemble's per-seed set-up
 * (examples/06_conditioned_fields/01_2D_condition_ensemble.py:32-35 draws one seed per realisation) stops being a
 * serial Python loop.  Per seed s, bit for bit what the reference draws for RandMeth(model, mode_no, seed=s) with a
 * model that has no inverse CDF (sampling through RNG.sample_ln_pdf) and a native log-pdf (pdf_kind):
 *   z_1, z_2 (mode_no,)  RandomState.normal                          generator.py:362-363
 *   ang_1, ang_2         uniform(0, two_pi) and, in 3-D, uniform(-1, 1) of RNG.sample_sphere (random/rng.py:163-174;
 *                        dim 1: the +-1 of choice([-1, 1])).  The caller finishes the sphere coordinates with numpy's
 *                        cos / sin / sqrt on the whole batch (numpy's SIMD trigonometry is not libm's).
 *   rad                  RNG.sample_ln_pdf(size=mode_no, sample_around)  random/rng.py:38-104
 * Outputs are (n_seeds, mode_no) row-major; seeds must fit 32 bits (numpy legacy seeding).  Host only.
 */
int gsb_sample_modes_batch(int pdf_kind, int dim, double len_rescaled, double nu, const int64_t *seeds,
                           int64_t n_seeds, int64_t mode_no, int nwalkers, int burn_in, int64_t n_steps,
                           double sample_around, double two_pi, int n_threads, double *z_1, double *z_2,
                           double *ang_1, double *ang_2, double *rad);

/*
 * Fused caller epilogue (reference: src/gstools/field/generator.py:269-270):
 *   field[i] = scale * field[i] + shift     in place, device pointers only.
 * Lets a device-resident caller keep the field on the GPU.
 */
int gsb_scale_shift(double *field, int64_t n, double scale, double shift, int device,
                    void *stream);

/*
 * Multi-GPU plan: ONE process drives several GPUs of a box (SURVEY.md section 8b "gsb_plan_create/destroy",
 * section 8e; the reference's callers are single-process: src/gstools/field/srf.py:150-163 calls the generator
 * once per field, examples/06_conditioned_fields/01_2D_condition_ensemble.py:32-35 loops over seeds).
 * Every output element depends only on its own position and the (tiny, replicated) mode set, so a call is cut
 * into independent shares with NO inter-GPU traffic during the sum:
 *     flat point sets          contiguous point ranges                              (gsb_plan_summate)
 *     meshes, n_batch <  n_dev slabs along axis 0 -- each device's output is one contiguous block of
 *                              every C-ordered field (tools/geometric.py:340-356)   (gsb_plan_summate_structured)
 *     meshes, n_batch >= n_dev contiguous ranges of batch entries (ensemble seeds)
 * Share g of n units over G devices is [lo, hi) with lo = g*(n/G) + min(g, n%G) (gsb_plan_share).
 * The plan owns one host thread per device; a call hands every device its share and returns when all are done.
 *   mem == GSB_MEM_HOST   : every device copies its share of the result straight into its slice of the caller's ONE
 *                           host array (pinned memory recommended), overlapped with its own compute.
 *   mem == GSB_MEM_DEVICE : all pointers live on `home_device` (a device of the plan).  The other devices read the
 *                           small inputs and STORE THEIR SHARE OF THE FIELD DIRECTLY INTO `out` on the home device
 *                           through NVLink peer memory from the contraction kernel's epilogue -- compute and
 *                           gather are one kernel, nothing is copied afterwards.  Work is ordered after `stream`
 *                           (a stream of the home device) and `stream` waits for all devices before later work.
 *                           Needs peer access between the plan's devices (gsb_plan_create enables it; the call
 *                           fails with GSB_ERR_ARGUMENT when the topology does not allow it).
 * `incompr` != 0 selects summate_incompr.  `epi` as in the *_ex entries.  `pepi` (scalar fields only): NULL, or an
 * array of n_devices per-device epilogues -- entry g holds gain / offset arrays resident on devices[g], each
 * covering ALL points of the field (device g reads its own share of them).
 * A plan is not re-entrant: one call at a time (calls from several threads are serialised).
 */
typedef struct gsb_plan gsb_plan;

/* devices == NULL or n_devices <= 0: all visible devices.  (A device may be listed more than once; its shares then
 * run one after the other.) */
int gsb_plan_create(const int *devices, int n_devices, gsb_plan **plan);
int gsb_plan_destroy(gsb_plan *plan);
/* Number of devices; devices (may be NULL) receives their indices; *peer_access: 1 when every pair has it. */
int gsb_plan_info(const gsb_plan *plan, int *n_devices, int *devices, int *peer_access);
/* The share [lo, hi) of part `part` of `parts` over n units (host only, no plan needed). */
int gsb_plan_share(int64_t n, int parts, int part, int64_t *lo, int64_t *hi);

int gsb_plan_summate(gsb_plan *plan, const double *cov_samples, const double *z_1, const double *z_2,
                     const double *pos, int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts, double *out,
                     int64_t out_ld, int incompr, const gsb_epilogue *epi, const gsb_point_epilogue *pepi,
                     int mem, int home_device, void *stream);

int gsb_plan_summate_structured(gsb_plan *plan, const double *cov_samples, const double *z_1, const double *z_2,
                                const double *axes, const int64_t *axis_len, const double *matrix, int dim,
                                int64_t n_modes, int64_t n_batch, double *out, int incompr,
                                const gsb_epilogue *epi, const gsb_point_epilogue *pepi, int mem,
                                int home_device, void *stream);

/*
 * gsb_plan_krige_evaluate -- gsb_krige_evaluate[_structured] (row f1: the chunk loop of Krige.__call__,
 * src/gstools/krige/base.py:278-294, with the right-hand sides of :359-388 generated on the device) over the devices
 * of a plan.  Points are independent and the kriging system is replicated: meshes (pos == NULL: `axes` / `axis_len` /
 * `matrix`) are cut into slabs along axis 0, flat point sets (`pos` (dim, n_pts), row stride pos_ld) into contiguous
 * ranges; drift rows follow their points.  mem / home_device / stream as for gsb_plan_summate_structured.
 */
int gsb_plan_krige_evaluate(gsb_plan *plan, const gsb_cov_model *model, const double *krig_mat, const double *cond,
                            int64_t krige_size, const double *cond_pos, int64_t cond_no, int dim, const double *pos,
                            int64_t pos_ld, int64_t n_pts, const double *axes, const int64_t *axis_len,
                            const double *matrix, int unbiased, const double *tail_rows, int64_t tail_ld,
                            double *field, double *error, int mem, int home_device, void *stream);

/*
 * gsb_summate_structured_slab -- one share of a mesh for callers that do their own fan-out (one process per GPU,
 * gstools_b200/dist.py): evaluates the entries [slab_lo, slab_hi) of axis 0 only, while `out` (and the per-point
 * arrays of `pepi`) describe the FULL mesh: the slab of field z lands at out + z*n + slab_lo*(n/axis_len[0]).
 * With GSB_MEM_DEVICE `out` may be memory of ANOTHER device mapped into this process (peer access or
 * gsb_ipc_open): the contraction kernel then stores its slab straight into the gathered field over NVLink, so
 * "sum, then gather onto rank 0"  is one kernel and no copy.
 */
int gsb_summate_structured_slab(const double *cov_samples, const double *z_1, const double *z_2, const double *axes,
                                const int64_t *axis_len, const double *matrix, int dim, int64_t n_modes,
                                int64_t n_batch, int64_t slab_lo, int64_t slab_hi, double *out, int incompr,
                                const gsb_epilogue *epi, const gsb_point_epilogue *pepi, int mem, int device,
                                void *stream);

/*
 * Device memory shared between the processes of one box (one process per GPU): the owner exports the allocation
 * that contains `dev_ptr` (64-byte handle + the pointer's offset inside the allocation), another process maps it
 * with gsb_ipc_open on its own `device` (peer access is enabled on demand) and gets the allocation's base address
 * in its address space; gsb_ipc_close unmaps.  Thin wrappers of cudaIpcGetMemHandle / OpenMemHandle / CloseMemHandle.
 */
#define GSB_IPC_HANDLE_BYTES 64
int gsb_ipc_export(const void *dev_ptr, int device, unsigned char *handle, int64_t *offset);
int gsb_ipc_open(const unsigned char *handle, int device, void **base);
int gsb_ipc_close(void *base, int device);

/*
 * Tuning / introspection (defaults in brackets; none of them changes results beyond the rounding of another tiling).
 *   gsb_set_option("force_path", 0|1|2): structured meshes: [0] cost model, 1 always direct, 2 always separable.
 *   gsb_set_option("structured_min_tiles", v): meshes with fewer 128x128 output tiles than v go to the direct kernel [0].
 *   gsb_set_option("fold_axes", 0|1|2): fold the last two axes of a thin mesh into one column axis of the structured
 *       path: 0 never, [1] when the cost model prefers it, 2 whenever it fits.
 *   gsb_set_option("sk_pack", 0|1|2): row tiles packed across slow indices: [0] cost model, 1 never, 2 whenever possible.
 *   gsb_set_option("sk_table_mb", v): cap (MiB) on the tile-axis table of a folded tile axis [256].
 *   gsb_set_option("sk_grid", v): CTAs of the persistent contraction, 0 = [one per SM] (tests force small / odd grids).
 *   gsb_set_option("host_pieces", n): host route: n equal pieces instead of [0] = growing pieces.
 *   gsb_set_option("host_chunk_points", v): host route of the direct path: points per chunk [4 Mi].
 *   gsb_set_option("direct_cfg", -1|0|1|2): direct kernel: [-1] cost model, else 1 / 2 / 8 points per thread.
 *   gsb_set_option("direct_split", 0|1): direct kernel: [1] small point sets split the mode loop over CTAs (same bits).
 *   gsb_set_option("scratch_mb", v): scratch budget (MiB) of the kriging right-hand sides generated on the device [3072].
 *   gsb_set_option("krige_host_chunk_mb", v): host route of the native-signature kriging entry: chunk size [256].
 *   gsb_set_option("time_kernels", 0|1): see gsb_kernel_times().   gsb_set_option("trace", 0|1): print the timeline of the
 *       contraction launches and D2H pieces of a structured call to stderr (single caller only).
 *   gsb_get_counter("launches"): CUDA kernels launched by this library so far (process-wide);
 *   "direct_calls" / "separable_calls" / "sk_calls" / "packed_calls" / "folded_calls" / "krige_calls": how often each path ran.
 */
int gsb_set_option(const char *name, int64_t value);

/*
 * Scratch (mode records, phase tables, the pre-tiled operands, staging buffers of the host routes)
 * comes from the device's stream-ordered memory pool and is kept there between calls so that steady-state
 * calls allocate nothing.  gsb_release_memory() waits for the device and returns the cached memory to the
 * driver (e.g. before handing the GPU to another library in the same process).
 */
int gsb_release_memory(int device);
int64_t gsb_get_counter(const char *name);

/*
 * Introspection of the structured path's work split (host only, no device needed): the separable contraction
 * is ONE persistent launch whose CTAs take equal-cost contiguous shares of the (output tile, pipeline stage)
 * iteration space ("stream-K", gstools_b200/csrc/gsb_sepk.cuh).  For the tiles [tile_begin, tile_end) of a mesh
 * with `ly` tile-axis rows and `lc` columns (tile = ((field * n_slow + slow) * ceil(ly/128) + row tile) *
 * ceil(lc/128) + column tile) and `n_stages` stages per tile, writes the share boundaries of `max_grid` CTAs:
 * CTA c works on (tiles[c], stages[c]) inclusive .. (tiles[c+1], stages[c+1]) exclusive.  tiles / stages hold
 * max_grid + 1 entries; *grid receives the number of CTAs actually used.
 */
int gsb_streamk_plan(int64_t tile_begin, int64_t tile_end, int64_t ly, int64_t lc, int n_stages,
                     int max_grid, int64_t *tiles, int32_t *stages, int *grid);

/*
 * Device-side timing of the dominant kernels (roofline evidence).  While the option
 * "time_kernels" is 1, CUDA events are recorded on the launch stream around every direct /
 * separable kernel launch.  gsb_kernel_times() waits for them, returns the summed duration and
 * the number of launches since the previous call, and clears the list.
 */
int gsb_kernel_times(double *total_ms, int64_t *n_launches);

/*
 * FP64-pipe microbenchmark used as the roofline denominator: runs a register-resident
 * DFMA loop on every SM and reports sustained DFMA/s (fused multiply-adds per second,
 * 1 DFMA = 2 flop).  `kind` 0 = DFMA (CUDA cores), 1 = DMMA m8n8k4 (tensor path).
 */
int gsb_measure_fp64_peak(int device, int kind, double seconds, double *fma_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GSB200_H */
