/* bnnp.h -- C ABI of the B200 (sm_100a) SG-MCMC sampler kernels.
 *
 * This is the drop-in boundary for ONE path of ratschlab/bnn_priors: the inner
 * loop of bnn_priors/mcmc (SGLD, VerletSGLD, HMC) plus the gradient and log-density
 * of the elementwise priors (Normal / Laplace / Student-t and the other closed forms
 * of bnn_priors/prior, incl. sampled scales).  Each entry point names the reference
 * code it replaces (paths relative to the reference checkout).  The per-epoch test
 * evaluation has its own header, bnnp_eval.h.
 *
 * Conventions
 *  - plain C types only; every pointer inside BnnpLaunch is a DEVICE pointer
 *    owned by the caller; the library never allocates, frees or synchronises.
 *  - `stream` is a cudaStream_t passed as void*; all work is stream-ordered.
 *  - return value 0 = success, negative = BNNP_E_* (bad argument), positive = a
 *    cudaError_t; bnnp_last_error() gives the text (thread-local).
 *  - no exceptions cross this boundary.
 *
 * Data layout (DESIGN.md "Layout in HBM").  One chain = flat fp32 arrays
 * P, G, M (parameters, p.grad, momentum_buffer) of `total` floats, optional
 * PREV_P, PREV_G, PREV_M (verlet_sgld.py:72-83 snapshots) and an optional
 * replay-noise array, all with the SAME layout: tensor t (a "segment") occupies
 * [off_t, off_t + numel_t), off_t a multiple of BNNP_SEG_ALIGN floats; the gap
 * up to the next segment is padding the kernels keep at zero.  A segment is cut
 * into chunks of BNNP_CHUNK floats; one CTA processes one chunk (BnnpChunk).
 */
#ifndef BNNP_H
#define BNNP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNNP_ABI_VERSION 10

#define BNNP_SEG_ALIGN 32      /* floats: every segment starts on a 128-byte line */
#ifndef BNNP_THREADS
#define BNNP_THREADS 256       /* threads per CTA                                 */
#endif
#ifndef BNNP_UNROLL
#define BNNP_UNROLL 4          /* 128-bit accesses per thread per stream          */
#endif
#define BNNP_CHUNK (BNNP_THREADS * BNNP_UNROLL * 4)   /* 4096 floats per CTA      */

#define BNNP_NRED 8            /* partial sums per chunk                          */
#define BNNP_STATE_STRIDE 16   /* doubles per segment in seg_state                */

enum { BNNP_E_ARG = -1, BNNP_E_ALIGN = -2, BNNP_E_UNSUPPORTED = -3 };

/* prior kinds: elementwise priors with constant hyper-parameters.  NONE = the prior
 * gradient is already inside G.  (prior_loc, prior_scale, prior_df) hold:
 *   NORMAL       loc, scale, -          prior/loc_scale.py:34-35
 *   LAPLACE      loc, scale, -          prior/loc_scale.py:66-67
 *   STUDENT_T    loc, scale, df         prior/loc_scale.py:74-77
 *   CAUCHY       loc, scale, -          prior/loc_scale.py:70-71
 *   GENNORM      loc, scale, beta       prior/loc_scale.py:80-83, distributions.py:75-79
 *   LOGNORMAL    loc, scale, -          prior/loc_scale.py:86-92 (density of p, incl. "- p")
 *   UNIFORM      low, high - low, -     prior/transformed.py:12-47 (constant density of p)
 *   IMPROPER     -, -, -                prior/loc_scale.py:94-97 (log_prob == 0)
 *   DOUBLE_GAMMA loc, scale, concentration   prior/transformed.py:83-96
 *
 * Hierarchical priors (prior/hierarchical.py, prior/empirical_bayes.py): the scale of a
 * NORMAL / LAPLACE / STUDENT_T segment may be a sampled scalar.  That scalar u is a segment
 * of its own (numel 1, a "hyper segment") whose kind says how the scale s follows from u and
 * which density s has; (prior_loc, prior_scale) hold its two constants (a, b); the two
 * segments name each other in BnnpSegment.link:
 *   HYPER_GAMMA      s = softplus(u),     s ~ Gamma(concentration a, rate b)
 *                    prior/transformed.py:50-63; NormalGamma, LaplaceGamma, StudentTGamma
 *   HYPER_UNIFORM    s = a + b Phi(u),    s ~ Uniform(a, a + b)
 *                    prior/transformed.py:12-47; NormalUniform, LaplaceUniform, StudentTUniform
 *   HYPER_HALFCAUCHY s = softplus(u) b,   s ~ HalfCauchy(scale a)
 *                    prior/transformed.py:66-80; Horseshoe
 *   HYPER_IMPROPER   s = softplus(u),     log density 0
 *                    prior/loc_scale.py:100-103; NormalEmpirical, LaplaceEmpirical */
enum {
    BNNP_PRIOR_NONE = 0, BNNP_PRIOR_NORMAL = 1, BNNP_PRIOR_LAPLACE = 2, BNNP_PRIOR_STUDENT_T = 3,
    BNNP_PRIOR_CAUCHY = 4, BNNP_PRIOR_GENNORM = 5, BNNP_PRIOR_LOGNORMAL = 6, BNNP_PRIOR_UNIFORM = 7,
    BNNP_PRIOR_IMPROPER = 8, BNNP_PRIOR_DOUBLE_GAMMA = 9,
    BNNP_PRIOR_HYPER_GAMMA = 10, BNNP_PRIOR_HYPER_UNIFORM = 11, BNNP_PRIOR_HYPER_HALFCAUCHY = 12,
    BNNP_PRIOR_HYPER_IMPROPER = 13
};

/* which sampler's bookkeeping the per-segment epilogue applies */
enum {
    BNNP_OP_SGLD = 0,            /* mcmc/sgld.py:119-154                          */
    BNNP_OP_VERLET = 1,          /* mcmc/verlet_sgld.py:149-197                   */
    BNNP_OP_HMC = 2,             /* mcmc/hmc.py:41-79                             */
    BNNP_OP_SAMPLE_MOMENTUM = 3, /* mcmc/sgld.py:57-69                            */
    BNNP_OP_REDUCE = 4           /* dot(g,g), dot(m,m), sum log-prior; no writes:
                                    sgld.py:9-11, verlet_sgld.py:44-47, hmc.py:32-33,
                                    models/base.py:25-30                          */
};
enum { BNNP_PHASE_INITIAL = 0, BNNP_PHASE_MID = 1, BNNP_PHASE_FINAL = 2 };

/* BnnpLaunch.flags */
enum {
    BNNP_F_READ_P = 1u << 0,
    BNNP_F_READ_G = 1u << 1,
    BNNP_F_READ_M = 1u << 2,
    BNNP_F_WRITE_P = 1u << 3,       /* p += (cp*M) * m'                              */
    BNNP_F_WRITE_M = 1u << 4,       /* momentum_buffer <- m'                         */
    BNNP_F_SAVE_STATE = 1u << 5,    /* verlet_sgld.py:72-83 fused into the step      */
    BNNP_F_CALC_METRICS = 1u << 6,  /* est_temperature / est_config_temp             */
    BNNP_F_LOG_PRIOR = 1u << 7,     /* also reduce sum log p(theta) at the theta left
                                       in P by this launch                           */
    BNNP_F_CLAMP_GRAD = 1u << 8,    /* inference.py:219-220 re-applied to g + prior  */
    BNNP_F_NOISE_FIRST = 1u << 9,   /* m' = (cn*eps + cg*M*g) + cm*m, the rounding
                                       order of verlet_sgld.py:163-167; otherwise
                                       m' = (cm*m + cg*M*g) + cn*eps, sgld.py:129,142 */
    BNNP_F_MM_PRE_NOISE = 1u << 10, /* sgld.py:132-137 (momentum == 0): the metric is
                                       dot(m', m') before the noise is added         */
    BNNP_F_UPDATE_SQ = 1u << 11,    /* square_avg moving average, sgld.py:153-154    */
    BNNP_F_PRIOR_GRAD = 1u << 12,   /* g <- g - (1/N) dlog p/dtheta in-register: models/
                                       base.py:72-77 + inference.py:218 without autograd */
    BNNP_F_ALL_SUMS = 1u << 13,     /* reduce every dot product even without CALC_METRICS
                                       (HMC initial/final need m.m: hmc.py:50-53,32-33);
                                       otherwise a launch only reduces what its epilogue
                                       reads: g.g always, g.m and g.m' for VERLET       */
    BNNP_F_HYPER_POST = 1u << 16,   /* a step launch (WRITE_P | PRIOR_GRAD | LOG_PRIOR, all chunks)
                                       of a chain whose sampled scales all belong to NORMAL /
                                       LAPLACE segments: its epilogue -- to be applied with
                                       bnnp_finalize before the next launch -- does what the
                                       BNNP_F_HYPER pre-pass would do for the parameters this
                                       launch leaves in P, because for those two densities the
                                       statistic of d log p / d scale follows from the sum of
                                       log-density terms the launch reduces anyway: for such
                                       segments the launch reduces sum(-d^2/2) resp. sum(-|d|),
                                       independent of the scale it used.  Saves the pre-pass of
                                       the NEXT step; with BNNP_F_HYPER_CHAIN on that next launch
                                       also the bnnp_finalize in between.                   */
    BNNP_F_HYPER_CHAIN = 1u << 17,  /* this launch (all chunks, not the pre-pass) carries a PENDING
                                       BNNP_F_HYPER_POST epilogue instead of having it finalised first:
                                       a segment with a sampled scale derives its scale, a hyper segment
                                       its gradient term, from the pending launch's partial records (the
                                       hyper-parameter that launch left is stashed there), while the CTAs
                                       that apply the epilogue write the same values into the segment
                                       table / state for the host.  A step with sampled scales is then ONE
                                       launch, like any other step                                   */
    BNNP_F_REVERSE = 1u << 15,      /* CTA i processes chunk nchunks-1-i.  A chain larger than
                                       L2 that alternates the direction from launch to launch
                                       starts each launch on the lines the previous one
                                       touched last, which are still in the 126 MB L2
                                       (results do not depend on the order)               */
    BNNP_F_HYPER = 1u << 14         /* the hierarchical-prior pre-pass (BNNP_OP_REDUCE with
                                       READ_P | LOG_PRIOR, no writes to P): linked segments
                                       take their scale from the hyper-parameter u now in P,
                                       the launch reduces sum log p and the statistic of
                                       d log p / d scale, and its epilogue -- which must be
                                       applied with bnnp_finalize before the next launch --
                                       leaves, per hyper segment, the new scale in the segment
                                       table, -(1/N) d log p / du in BNNP_S_HYPER and the
                                       hyper log-density in BNNP_S_LOG_PRIOR.  Replaces
                                       autograd through scale_prior() (prior/base.py:52-58,
                                       prior/transformed.py:57-63)                      */
};

/* noise source */
enum { BNNP_NOISE_NONE = 0, BNNP_NOISE_REPLAY = 1, BNNP_NOISE_PHILOX = 2 };

/* indices into one segment's BNNP_STATE_STRIDE doubles */
enum {
    BNNP_S_DELTA_ENERGY = 0,   /* state['delta_energy']                       */
    BNNP_S_PREV_NEW_MOM = 1,   /* state['prev_new_momentum_delta']            */
    BNNP_S_EST_MM = 2,         /* numerator of state['est_temperature']       */
    BNNP_S_EST_PG = 3,         /* dot(p, grad) of state['est_config_temp']    */
    BNNP_S_SUM_GG = 4,         /* dot(grad, grad) seen by the last launch     */
    BNNP_S_SUM_MM = 5,         /* dot(m, m) of the momentum now in M          */
    BNNP_S_SQ_MEAN = 6,        /* mean(state['square_avg'])                   */
    BNNP_S_LOG_PRIOR = 7,      /* sum log p(theta)  (BNNP_F_LOG_PRIOR)        */
    BNNP_S_GM_OLD = 8,         /* raw sums of the last launch ...             */
    BNNP_S_GM_NEW = 9,
    BNNP_S_MM_OLD = 10,
    BNNP_S_MM_NEW = 11,
    BNNP_S_NONFINITE = 12,     /* 1.0 if the gradient had a non-finite entry  */
    BNNP_S_LAUNCHES = 13,      /* number of launches that finalised this segment */
    BNNP_S_HYPER = 14          /* hyper segment: -(1/N) d log p / du, added to its gradient
                                  by launches with BNNP_F_PRIOR_GRAD; linked weight segment:
                                  the statistic of d log p / d scale (sum d^2, sum |d|, sum
                                  d^2 / (df s^2 + d^2))                                    */
};

/* One parameter tensor.  A table of these lives in device memory. */
typedef struct BnnpSegment {
    int64_t off;         /* element offset in the flat layout (multiple of
                            BNNP_SEG_ALIGN)                                     */
    int64_t numel;       /* > 0                                                */
    double precond;      /* state['preconditioner'] (sgld.py:47-52)            */
    float prior_loc, prior_scale, prior_df;
    int32_t prior_kind;
    int32_t first_chunk; /* index of this segment's first chunk                */
    int32_t num_chunks;  /* ceil(numel / BNNP_CHUNK)                           */
    int32_t link;        /* hierarchical priors: index of the hyper segment that holds
                            this segment's scale, or (for a hyper segment) of the
                            segment it scales; -1: none                         */
    int32_t reserved;
} BnnpSegment;

/* One chunk of BNNP_CHUNK floats (the work of one CTA), 16 bytes: what a CTA needs before it
 * can issue its loads, in one 128-bit access (the segment descriptor is fetched in parallel
 * with the data instead of in front of it). */
typedef struct BnnpChunk {
    int64_t fbase;       /* flat index of the chunk's first float                */
    int32_t rem;         /* valid floats in this chunk (BNNP_CHUNK except at the
                            end of a segment)                                   */
    int32_t seg;         /* segment the chunk belongs to                        */
} BnnpChunk;

/* What the per-segment scalar bookkeeping (the "epilogue") of a launch needs.  It is
 * not applied by the launch itself but by the NEXT bnnp_launch on the same chain
 * (which receives it as BnnpLaunch.pending; its first nseg CTAs do the work) or by
 * bnnp_finalize.  A launch with chunk_ids must not carry a pending epilogue. */
typedef struct BnnpEpilogue {
    int32_t valid;           /* 0: nothing pending                                  */
    int32_t op, phase;
    uint32_t flags;
    int32_t parity;          /* which half of `partials` / `stamps` that launch wrote */
    int32_t reserved;
    uint64_t call;           /* its launch counter (stamps hold call + 1)           */
    double c_gm_base, curv_base, rms_alpha;
    double inv_num_data;     /* 1/N (BNNP_F_HYPER)                                   */
} BnnpEpilogue;

/* Coefficients of one kind of launch, as they sit in a device control block (BnnpControl). */
typedef struct BnnpCoef {
    double cm, cg, cn, cp;
    double inv_num_data;
    double c_gm_base, curv_base, rms_alpha;
} BnnpCoef;

#define BNNP_COEF_SLOTS 4      /* by convention: the sampler's PHASE_INITIAL / MID / FINAL transition
                                  and sample_momentum                                  */

/* Device-resident launch state of one chain ("capturable" mode).  A launch whose BnnpLaunch.ctl
 * points here takes its Philox counter, its parity (and with it the walking direction,
 * BNNP_F_REVERSE), the pending epilogue and its coefficients (slot BnnpLaunch.coef_slot) from
 * this block instead of from the kernel parameters, and bnnp_advance -- enqueued right after it --
 * moves the block on: pending <- this launch's epilogue, call += 1, parity ^= 1.  Nothing per-launch
 * is baked into the kernel parameters any more, so a step can be captured in a CUDA graph and
 * replayed: every replay draws new noise, alternates the direction and folds the previous replay's
 * sums, and a learning-rate / temperature change between replays is a bnnp_poke of the coefficient
 * slots (sgld.py:114-117, verlet_sgld.py:96-146 recomputed on the host as before).             */
typedef struct BnnpControl {
    uint64_t call;             /* Philox counter words 2,3 of the NEXT launch             */
    int32_t parity;            /* half of partials / stamps the NEXT launch writes         */
    int32_t reserved;
    BnnpEpilogue pending;      /* epilogue of the LAST launch (valid = 0: none)            */
    BnnpCoef coef[BNNP_COEF_SLOTS];
} BnnpControl;

typedef struct BnnpLaunch {
    float* P;                  /* flat [total]                                      */
    float* G;                  /* flat [total]: p.grad of every tensor, unless seg_grad
                                  is given; always the target of bnnp_rollback        */
    float* M;
    float* prev_p;             /* flat [total] or null (BNNP_F_SAVE_STATE)          */
    float* prev_g;
    float* prev_m;
    const float* replay_noise; /* flat [total], BNNP_NOISE_REPLAY                   */
    BnnpSegment* segs;         /* [nseg]; only the epilogue of a BNNP_F_HYPER launch
                                  writes it (prior_scale of linked segments)        */
    const BnnpChunk* chunks;   /* [all chunks], 16-byte aligned (bnnp_plan_layout)  */
    const int32_t* chunk_ids;  /* null: process chunks 0..nchunks-1; else [nchunks]
                                  chunk indices, whole segments only (used to skip
                                  tensors without a gradient, sgld.py:96-101)       */
    const float* const* seg_grad;  /* null, or device array [nseg]: the gradient of segment s is
                                  the contiguous, 16-byte aligned fp32 array seg_grad[s]
                                  (numel_s floats) instead of G + off_s -- the tensors
                                  autograd hands over after zero_grad() are read in place,
                                  no copy into G (sgld.py:94-105 reads p.grad likewise)  */
    BnnpControl* ctl;          /* null, or the chain's device control block: call, parity,
                                  BNNP_F_REVERSE, pending and the coefficients cm..rms_alpha
                                  of this struct are ignored and read from *ctl           */
    double* seg_state;         /* [nseg][BNNP_STATE_STRIDE]                         */
    double* partials;          /* scratch [2][nchunks_total][BNNP_NRED]             */
    uint64_t* stamps;          /* [2][nchunks_total], zero-initialised once         */
    int32_t nseg, nchunks;     /* nchunks: CTAs of this launch (= nchunks_total
                                  unless chunk_ids is given)                        */
    int32_t nchunks_total;     /* chunks of the whole chain (bnnp_plan_layout)      */
    int32_t parity;            /* 0/1: half of partials/stamps this launch writes;
                                  must differ from pending.parity                   */
    int32_t op, phase, noise;
    uint32_t flags;
    int32_t coef_slot;         /* which BnnpControl.coef entry this launch uses (ctl != null) */
    int32_t reserved;
    uint32_t key0, key1;       /* Philox key                                        */
    uint64_t call;             /* Philox counter words 2,3: one value per launch    */
    /* m' from cm*m, (cg*M)*g, cn*eps ;  p' = p + (cp*M)*m'   (M = precond)         */
    double cm, cg, cn, cp;
    double inv_num_data;       /* 1/N for the fused prior gradient                  */
    double grad_max;           /* BNNP_F_CLAMP_GRAD                                 */
    double c_gm_base;          /* -bhn/2        (verlet_sgld.py:170)                */
    double curv_base;          /* N^2 b^2h^2/8  (verlet_sgld.py:46)                 */
    double rms_alpha;          /* sgld.py:153                                       */
    BnnpEpilogue pending;      /* epilogue of the previous launch on this chain, to
                                  be applied by this one (valid = 0: none)          */
} BnnpLaunch;

int bnnp_abi_version(void);
const char* bnnp_last_error(void);

/* SM count and L2 size of a device (grid sizing, bench bookkeeping). */
int bnnp_device_info(int device, int* sm_count, int* l2_bytes);
/* Resident CTAs per SM of one step-kernel instantiation: noise kind, prior fused or
 * not, BNNP_F_NOISE_FIRST or not, sums = 0 (g.g only) / 1 (+ g.m, g.m') / 2 (all). */
int bnnp_max_ctas_per_sm(int noise, int has_prior, int noise_first, int sums, int* out);

/* Host helper (no GPU needed): flat-layout offsets and chunk table for `nseg`
 * tensors.  Fills off[nseg], first_chunk[nseg], num_chunks[nseg]; returns the
 * totals.  chunks may be null; otherwise it must hold *total_chunks records
 * (call once with null to size it). */
int bnnp_plan_layout(const int64_t* numel, int nseg,
                     int64_t* off, int32_t* first_chunk, int32_t* num_chunks,
                     int64_t* total_elems, int32_t* total_chunks, BnnpChunk* chunks);

/* The one hot kernel.  Replaces, depending on op/phase/flags:
 *   SGLD._step_fn            mcmc/sgld.py:119-154
 *   VerletSGLD._step_fn      mcmc/verlet_sgld.py:149-197 (+ _save_state :72-83,
 *                            _point_energy :44-47)
 *   HMC._step_fn             mcmc/hmc.py:41-79
 *   SGLD.sample_momentum     mcmc/sgld.py:57-69
 *   dot()                    mcmc/sgld.py:9-11
 *   Prior.log_prob + autograd for Normal/Laplace/StudentT
 *                            prior/base.py:57-58, prior/loc_scale.py:34-77,
 *                            models/base.py:72-77, inference.py:218-220        */
int bnnp_launch(const BnnpLaunch* args, void* stream);

/* Apply args->pending (the scalar bookkeeping of the last launch: state['delta_energy'],
 * ['prev_new_momentum_delta'], est_* numerators, square_avg mean, log-prior --
 * verlet_sgld.py:170-187, hmc.py:50-72, sgld.py:127-154) now, in a launch of nseg small
 * CTAs.  Needed before the host reads seg_state, changes the segment table, or starts
 * a launch that skips segments.  Reads segs, seg_state, partials, stamps, nseg,
 * nchunks_total and pending; no-op if pending.valid == 0. */
int bnnp_finalize(const BnnpLaunch* args, void* stream);

/* Capturable mode: move the control block on after a bnnp_launch with args->ctl != null (same
 * args): ctl->pending <- that launch's epilogue (op, phase, flags from args; parity, call and the
 * epilogue coefficients from *ctl), ctl->call += 1, ctl->parity ^= 1.  One tiny launch. */
int bnnp_advance(const BnnpLaunch* args, void* stream);

/* Capturable mode: bnnp_finalize applies ctl->pending when args->ctl != null; this clears it
 * afterwards (ctl->pending.valid = 0). */
int bnnp_clear_pending(BnnpControl* ctl, void* stream);

/* Stream-ordered write of a small HOST buffer into device memory, carried in the parameters of a
 * tiny kernel (nbytes <= 3840, a multiple of 4; dst 4-byte aligned): no pinned staging buffer whose
 * lifetime the caller would have to manage, legal during stream capture.  Used for the per-segment
 * gradient pointers (BnnpLaunch.seg_grad) when autograd moved a gradient, and for BnnpControl
 * (initial state, coefficient slots). */
int bnnp_poke(void* dst, const void* src_host, int64_t nbytes, void* stream);

/* VerletSGLD.maybe_reject's restore (mcmc/verlet_sgld.py:63-69):
 * P,G,M <- prev_*  over `total` floats (prev_m/M may be null: momentum == 0). */
int bnnp_rollback(float* P, float* G, float* M, const float* prev_p, const float* prev_g,
                  const float* prev_m, int64_t total, void* stream);

/* Diagnostic only (tools/tune_tiles.py): the bare memory access pattern of a step
 * (read P, G, M; write P, M over `total` floats, two FMAs per element) -- the
 * bandwidth ceiling the step kernel is compared against.  Overwrites P and M. */
int bnnp_probe_stream(float* P, const float* G, float* M, int64_t total, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BNNP_H */
