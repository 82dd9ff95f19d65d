/*
 * reheatfunq_b200 — C ABI of the B200-native anomaly-posterior backend.
 *
 * This is the drop-in boundary for the hot path of mjziebarth/REHEATFUNQ.  The
 * reference has no C ABI: its Cython modules bind C++ classes directly
 * (`cdef extern` in reheatfunq/anomaly/postbackend.pyx:38-120).  Each entry
 * point below names the reference interface it replaces.  Conventions:
 *   - plain pointers and sizes only, all arrays caller-owned and contiguous;
 *   - ragged data in CSR form (values + offsets);
 *   - every function returns an int status (0 = OK, see RHFQ_ST_*) and writes a
 *     NUL-terminated message into (err, errlen) on failure — the Python shim
 *     re-raises it as RuntimeError, the exception type the reference's
 *     `except+` translation produces (postbackend.pyx:98-120);
 *   - a handle is immutable after creation except for the lazily built Simpson
 *     tree (as in the reference, posterior.hpp:155-156); calls on ONE handle may come from
 *     several host threads and are serialised inside the library;
 *   - every entry point restores the caller's current CUDA device before it returns;
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     RHFQ_ST_NO_DEVICE.
 */
#ifndef REHEATFUNQ_B200_H
#define REHEATFUNQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes */
#define RHFQ_ST_OK 0
#define RHFQ_ST_RUNTIME 1            /* NaN/inf result, internal failure                  */
#define RHFQ_ST_DOMAIN 2             /* q <= 0, Qmax == inf, duplicate Qmax, bad weight,
                                        quantile outside [0,1]  (locals.hpp:228-252,
                                        posterior.hpp:539-541,313-317)                    */
#define RHFQ_ST_NOT_NORMALIZABLE 3   /* n < v or degenerate n == v (integrand.hpp:228-237)*/
#define RHFQ_ST_PRECISION 4          /* quadrature tolerance not reached
                                        (integrand.hpp:459-466)                           */
#define RHFQ_ST_ROOT 5               /* bracketed root search failed (large_z.hpp:345-346)*/
#define RHFQ_ST_ARGUMENT 6           /* invalid argument to this API                      */
#define RHFQ_ST_NO_DEVICE 7          /* no CUDA device / CUDA runtime failure             */

/* pdf algorithms (include/anomaly/posterior.hpp:78-82) */
#define RHFQ_ALG_EXPLICIT 0
#define RHFQ_ALG_BARYCENTRIC_LAGRANGE 1
#define RHFQ_ALG_ADAPTIVE_SIMPSON 2

/* flags for rhfq_batch_create */
#define RHFQ_FLAG_INPUTS_ON_DEVICE 1 /* q, c, w, offsets, prior are device pointers       */

typedef struct rhfq_batch rhfq_batch;

const char* rhfq_version(void);
int rhfq_device_count(void);

/*
 * Creates R independent posteriors in one call (the batched form of
 * VariablePrecisionPosterior's constructor, include/anomaly/
 * variableprecisionposterior.hpp:243-252; what CppAnomalyPosterior.__init__
 * does for R = 1, postbackend.pyx:170-246).
 *
 *   q, c       heat flow and anomaly factors of all sample sets, concatenated
 *   set_off    [S+1] offsets of the S sample sets into q / c
 *   w          [S]   sample-set weights (posterior.hpp:70-76)
 *   post_off   [R+1] offsets of the R posteriors into the sets
 *   n_sets, n_data   lengths of w (= S) and of q / c: both offset arrays must start at 0, be
 *              non-decreasing and end at these lengths, else RHFQ_ST_ARGUMENT is returned
 *              before anything is read through them
 *   prior      (p, s, n, v): 4 doubles shared by all posteriors (prior_stride 0)
 *              or 4 per posterior (prior_stride 4)
 *   algorithm  RHFQ_ALG_*;  bli_max_refinements as in posterior.hpp:91
 *   device     CUDA device ordinal
 *   flags      bit 0: q, c, set_off, w, post_off, prior are device pointers;
 *              bit 1: the reference's LONG DOUBLE build is the model -- what its Python keyword
 *              precision="double" runs (postbackend.pyx:219-222).  The arithmetic stays IEEE
 *              double; the thresholds that are powers of the machine epsilon of the reference's
 *              type take their long double values: y_max criterion (large_z.hpp:314), tanh-sinh
 *              tolerance of the norm (localsandnorm.hpp:210), Simpson tolerance
 *              (simpson.hpp:222), PointInInterval threshold (intervall.hpp:58)
 * Failures of individual posteriors (domain, not normalisable, ...) do not
 * fail the call: query rhfq_batch_status().  For R == 1 the Python shim turns
 * a non-zero posterior status into the RuntimeError the reference throws.
 */
int rhfq_batch_create(rhfq_batch** out,
                      const double* q, const double* c, const int64_t* set_off,
                      const double* w, const int64_t* post_off, int64_t R,
                      int64_t n_sets, int64_t n_data,
                      const double* prior, int prior_stride,
                      double amin, double rtol, int algorithm, int bli_max_refinements,
                      int device, int flags, char* err, size_t errlen);

void rhfq_batch_destroy(rhfq_batch* b);
/* The large scratch buffers (Simpson node pool, node plans, interpolant samples: several GB for
 * 10^4 posteriors) of the most recently destroyed batch are kept per device and adopted by the
 * next rhfq_batch_create there, and so is its CUDA stream; this call frees them (device < 0: all
 * devices).  RHFQ_NO_CACHE=1 in the environment disables the buffer cache. */
void rhfq_cache_release(int device);

/* Number of posteriors / Qmax per posterior (Posterior::get_Qmax, posterior.hpp:110-112) */
int64_t rhfq_batch_size(const rhfq_batch* b);
int rhfq_batch_qmax(const rhfq_batch* b, double* qmax /* [R] */);
/* Per-posterior status after creation (RHFQ_ST_*) */
int rhfq_batch_status(const rhfq_batch* b, int32_t* status /* [R] */);
const char* rhfq_status_message(int status);

/*
 * Evaluation.  Posterior r is evaluated at x[x_off[r] .. x_off[r+1]); results
 * go to out at the same positions.  Replaces Posterior::pdf / cdf / tail /
 * tail_quantiles (posterior.hpp:117-337) and the *_inplace forwards
 * (variableprecisionposterior.hpp:254-300, postbackend.pyx:273-366).
 * n_points: length of x and out; x_off must start at 0, be non-decreasing and end at
 * n_points (else RHFQ_ST_ARGUMENT, nothing is read or written).
 * on_device != 0: x, x_off and out are device pointers.
 * qstatus (may be NULL): [R] per-posterior status of this call.
 * Thread safety: calls on one handle may come from several host threads (the reference
 * releases the GIL around evaluation, postbackend.pyx:284-288); they are serialised inside.
 */
int rhfq_batch_pdf(rhfq_batch* b, const double* x, const int64_t* x_off, int64_t n_points,
                   double* out, int on_device, int32_t* qstatus, char* err, size_t errlen);
int rhfq_batch_cdf(rhfq_batch* b, const double* x, const int64_t* x_off, int64_t n_points,
                   double* out, int on_device, int32_t* qstatus, char* err, size_t errlen);
int rhfq_batch_tail(rhfq_batch* b, const double* x, const int64_t* x_off, int64_t n_points,
                    double* out, int on_device, int32_t* qstatus, char* err, size_t errlen);
int rhfq_batch_tail_quantiles(rhfq_batch* b, const double* t, const int64_t* t_off,
                              int64_t n_points, double* out, int on_device, int32_t* qstatus,
                              char* err, size_t errlen);

/*
 * Debug getters (Posterior::get_locals, posterior.hpp:381-409, and
 * get_log_pdf_bli_samples, :411-421; postbackend.pyx:368-455).
 * locals[36]: lp ls n v amin Qmax h0 h1 h2 h3 w lh0 l1p_w lv log_scale ymax ztrans S taylor
 *             norm weight log_fac N imax status k_off flags
 *             c1[0..1] c2[0..2] c3[0..3]  -- polynomial coefficients in a of the large-z
 *             constants C_1..C_3 (large_z.hpp:82-176; Posterior::get_C, posterior.hpp:427-440)
 * flags bit 0: the z-quadrature of the set's norm ran out of levels; like the reference
 * (localsandnorm.hpp:221-226, no check of the error estimate) the last estimate is used.
 */
int rhfq_batch_get_locals(const rhfq_batch* b, int64_t set_index, double* locals);
/* k_i of a sample set (Locals::ki, locals.hpp:75-87): returns its N (or < 0), writes min(N, cap) */
int64_t rhfq_batch_get_k(const rhfq_batch* b, int64_t set_index, double* k, int64_t cap);
/* which: 0 low / 1 high interpolant of posterior r.  Returns the number of kept
 * samples (or < 0); writes min(count, cap) values of f (log pdf) at the nodes. */
int64_t rhfq_batch_get_bli(const rhfq_batch* b, int64_t r, int which, double* f, int64_t cap);
/* Number of Simpson leaves of the whole batch (0 before the first cdf/tail call) */
int64_t rhfq_batch_saq_leaves(const rhfq_batch* b);

/*
 * Work counters since creation (the roofline's algorithmic work):
 * counters[21] = regular node evaluations, large-z node evaluations, evaluations
 * of the alpha log-integrand, Newton iterations, sum of N over regular nodes,
 * kernel launches, device time [ns] and launch count of the node-evaluation
 * kernel (CUDA events on the launching stream), the same two for the
 * norm-quadrature node kernel, Chebyshev terms and frontier intervals of the
 * Simpson-tree kernel, its device time [ns] and launch count, interpolant
 * evaluations through the fine theta grid, interpolants that failed its probe check,
 * device time [ns] and launch count of the mode/window (phase A) node kernels -- the node
 * evaluation kernel timed above is then the Gauss-Legendre (phase B) kernel -- and the
 * evaluations of the alpha log-integrand made by the phase B kernels, and the number of sample
 * sets whose z-quadrature ran out of levels (see rhfq_batch_get_locals), and the number of
 * nodes whose alpha-integral was re-evaluated by the sampled precision check: one node in 127
 * gets its Gauss-Legendre sums repeated with every panel cut in two; a relative difference above 1e-5
 * -- the threshold at which the reference throws PrecisionError (integrand.hpp:459-466) -- or a
 * window search that ran out of iterations gives the posterior RHFQ_ST_PRECISION.
 * The device times of the node / norm / mode-window kernels are recorded only for batches created
 * with RHFQ_KERNEL_TIMERS=1 in the environment (event pairs around ~60 launches cost 4-6 us of
 * device idle time each) and are 0 otherwise; launch counts and the Simpson kernel's time always are.
 */
int rhfq_batch_counters(const rhfq_batch* b, uint64_t* counters);

/*
 * Resilience Monte Carlo for ONE sample size N: replaces
 * test_performance_{1q,41q,mixture_4q,mixture_41q}
 * (include/resilience/resilience.hpp:38-69, src/resilience/resilience.cpp:511-575;
 * bound by reheatfunq/resilience/zeal2022hfresil.pyx:40-100).
 *
 *   dist_kind    0: gamma (params K, T);  1: two-normal mixture (x0, s0, a0, x1, s1, a1)
 *   N, M         sample size, number of realisations of the whole experiment
 *   shard_rank, shard_world   THIS call computes the realisations of the RNG streams
 *                t = shard_rank, shard_rank + shard_world, ... (multi-GPU: one call per rank;
 *                no rank replays another rank's draws and the experiment does not depend
 *                on shard_world).  (0, 1): the whole experiment.
 *   quantiles    [Nq] tail quantiles (any Nq >= 1; the reference allows 1/41 resp. 4/41)
 *   prior        (p, s, n, v) of the informed prior; the flat prior (1,0,0,0) is added
 *   tol          rtol of the posterior (zeal2022hfresil.pyx:111,227)
 *   seed, nstreams   std::seed_seq{seed} -> one mt19937_64 per stream; stream t owns the
 *                batches j = t, t + nstreams, ... of 10 realisations (for nstreams = 1
 *                identical to the reference, which is only reproducible in that case)
 *   chunk        realisations per device batch (0: default)
 *   out          [2][Nq][M_local], M_local = rhfq_resilience_shard_count(...): the shard's
 *                columns of res[:, i, :, :] (zeal2022hfresil.pyx:318-324) in increasing order
 *                of the global realisation index (rhfq_resilience_shard_index gives it; for
 *                shard_world = 1 local = global); NaN rows for failed posteriors.  Host
 *                memory, or device memory if out_on_device (the NCCL gather reads it there).
 *   failures     [2] number of failed informed / flat posteriors of the shard
 *   stats        [9] (may be NULL): regular nodes, large-z nodes, g evaluations, Newton
 *                iterations, sum N, launches, node-kernel ns, host draw time ns, sets whose
 *                z-quadrature ran out of levels
 *   dump_q, dump_c, dump_u, dump_q0   optional [M_local][N] dumps of the synthetic data (q, c
 *                as handed to the posterior; position uniforms; heat flow before the anomaly)
 *   dump_samples optional [M_local][4]: samples kept by the barycentric interpolants (informed
 *                low / high, flat low / high; 0 for a failed posterior) -- the refinement
 *                decisions (barylagrint.hpp:576-668) the parity tools compare with the oracle's
 */
int rhfq_resilience_run(int dist_kind, const double* dist_params, int64_t N, int64_t M,
                        int shard_rank, int shard_world, double P_MW, const double* quantiles,
                        int Nq, const double* prior, double amin, double tol, uint64_t seed,
                        int nstreams, int device, int64_t chunk, double* out, int out_on_device,
                        int64_t* failures, uint64_t* stats, double* dump_q, double* dump_c,
                        double* dump_u, double* dump_q0, int32_t* dump_samples, char* err,
                        size_t errlen);
/* Number of realisations of the shard, and their global indices index[M_local] (increasing):
 * stream t owns the batches j = t, t + nstreams, ... of 10 realisations
 * (src/resilience/resilience.cpp:397,420-437 deals the same batches out dynamically). */
int64_t rhfq_resilience_shard_count(int64_t M, int nstreams, int shard_rank, int shard_world);
int rhfq_resilience_shard_index(int64_t M, int nstreams, int shard_rank, int shard_world,
                                int64_t* index);

/*
 * Known-answer entry of the barycentric Lagrange interpolator: replaces
 * reheatfunq._testing.barylagrint (src/testing/barylagrint.cpp:32-145,
 * testing/test_numerics.py:28-95).  Runs the engine's refinement check, coefficient and
 * evaluation kernels on one of the reference test's functions over [xmin, xmax]:
 *   kind 0: exp(-x)   1: sin^2(x)   2: Heaviside(x - 2)   3: exp(-(x-2)^2 / (2 * 0.5^2))
 * with the constructor arguments of PiecewiseBarycentricLagrangeInterpolator
 * (barylagrint.hpp:107-112) and evaluates it at x[n].  barycentric != 0 evaluates with the
 * barycentric formula everywhere (default: Clenshaw on the Chebyshev coefficients).
 * info[2] (may be NULL): samples kept, kernel launches.
 */
int rhfq_bli_known_answer(int kind, double xmin, double xmax, double tol_rel, double tol_abs,
                          double fmin, double fmax, int max_refinements, const double* x,
                          int64_t n, double* out, int barycentric, int64_t* info, int device,
                          char* err, size_t errlen);

/*
 * Gamma conjugate prior integrals (external/pdtoolbox/include/
 * gamma_conjugate_prior.hpp:39-162; bound by reheatfunq/regional/backend.pyx:
 * gcp_ln_Phi :271-272, gamma_conjugate_prior_kullback_leibler :497-511,
 * gamma_conjugate_prior_kullback_leibler_batch_max :514-528, GCPMSECache :563-734).
 *
 * params / refs are rows (lp, s, n, v).
 * rhfq_gcp_ln_phi:       ln Phi of K tuples; status[K] (may be NULL).
 * rhfq_gcp_kl_batch_max: for each of the K candidate reference tuples the maximum
 *   over the N sample tuples of KL(sample_i || ref_k)  (ln Phi, the KL integral
 *   and the max of kullback_leibler_batch_max, gamma_conjugate_prior.cpp:667-712);
 *   +inf where the reference returns +inf (failed normalisation).  kl_all (may be
 *   NULL): [K][N] individual distances.  K = N = 1 is the scalar
 *   GammaConjugatePriorBase::kullback_leibler (:646-664).
 */
int rhfq_gcp_ln_phi(const double* params, int64_t K, double amin, double* out,
                    int32_t* status, int device, char* err, size_t errlen);
int rhfq_gcp_kl_batch_max(const double* params, int64_t N, const double* refs, int64_t K,
                          double amin, double* out, double* kl_all, int device,
                          char* err, size_t errlen);

/*
 * d_min-restricted sample enumeration (host only; src/coverings/dminpermutations.cpp:
 * determine_restricted_samples :525-679, global_PHmax :682-879; bound by
 * reheatfunq/coverings/rdisks.pyx: all_restricted_samples :324-353,
 * bootstrap_data_selection :229-318, determine_global_PHmax :358-369).
 *
 * xy is [N][2] (projected metres).  numpy_bitgen is the `bitgen_t*` of a numpy
 * BitGenerator (`bit_generator.ctypes.bit_generator`, numpy/random/bitgen.h) and
 * may be NULL for rhfq_restricted_samples (then exceeding max_samples/max_iter is
 * an error, like the reference without a generator).  The result is a handle
 * holding S samples: weights w[S], offsets off[S+1] and concatenated ascending
 * node indices idx[off[S]].
 */
typedef struct rhfq_samples rhfq_samples;
int rhfq_restricted_samples(rhfq_samples** out, const double* xy, int64_t N, double dmin,
                            int64_t max_samples, int64_t max_iter, void* numpy_bitgen,
                            char* err, size_t errlen);
int rhfq_bootstrap_data_selection(rhfq_samples** out, const double* xy, int64_t N, double dmin,
                                  int64_t B, void* numpy_bitgen, char* err, size_t errlen);
int64_t rhfq_samples_count(const rhfq_samples* s);
int64_t rhfq_samples_index_count(const rhfq_samples* s);
int rhfq_samples_get(const rhfq_samples* s, double* w, int64_t* off, int64_t* idx);
void rhfq_samples_destroy(rhfq_samples* s);
int rhfq_global_phmax(const double* xy, const double* phmax_i, int64_t N, double dmin,
                      double* out, char* err, size_t errlen);

/*
 * Sustained FP64 FMA throughput of the device in TFLOP/s, measured with a
 * register-resident DFMA chain (the roofline denominator: MEASURED_PEAKS.json
 * holds no FP64 figure).
 */
int rhfq_measure_fp64_peak(int device, double* tflops, char* err, size_t errlen);

#ifdef __cplusplus
}
#endif

#endif
