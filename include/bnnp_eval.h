/* bnnp_eval.h -- C ABI of the device-side test-set evaluation of posterior samples
 * (SURVEY 8f N3).  Part of libbnnp.so; same conventions as bnnp.h (plain C types,
 * device pointers owned by the caller, stream-ordered, 0 / BNNP_E_* / cudaError_t).
 *
 * Replaces the bookkeeping of evaluate_model (bnn_priors/exp_utils.py:250-340), which
 * the runners call once per epoch (inference.py:199-213): the reference moves every
 * batch's log-probabilities and logits to the CPU as float64 (one blocking copy per
 * batch, exp_utils.py:295-297), keeps an [E, N, C] float64 tensor there and reduces it
 * at the end (:301-322).  Here the per-point accumulators stay in HBM, are updated by
 * one launch per batch and reduced by one launch at the end; the host reads back five
 * doubles.  The network's forward pass stays the model's own torch code.
 */
#ifndef BNNP_EVAL_H
#define BNNP_EVAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    BNNP_EVAL_CATEGORICAL = 0,   /* preds is td.Categorical: exp_utils.py:281-283 */
    BNNP_EVAL_NORMAL = 1         /* preds is td.Normal:      exp_utils.py:284-286 */
};

/* indices into the `out` array of bnnp_eval_finish */
enum {
    BNNP_EV_LP_ENSEMBLE = 0,     /* exp_utils.py:305  _log_space_mean(lps, 0).mean()                 */
    BNNP_EV_LP_LAST = 1,         /* exp_utils.py:304  lps.mean(1)[-1]                                */
    BNNP_EV_ACC_ENSEMBLE = 2,    /* exp_utils.py:320  acc_mse(ensemble_preds, labels).mean(0)        */
    BNNP_EV_ACC_LAST = 3,        /* exp_utils.py:321  acc_mse(last_preds, labels).mean(0)            */
    BNNP_EV_LP_ENSEMBLE_CHECK = 4, /* exp_utils.py:312 ensemble_preds.log_prob(labels).mean(0)
                                      (categorical only; the reference asserts it equals [0])        */
    BNNP_EV_OUT = 8
};

#define BNNP_EVAL_ROW 5

/* Per-point accumulators of one evaluation, all [N] or [N][C] doubles in device memory. */
typedef struct BnnpEvalState {
    double* ens;        /* [N][C]  categorical: log-sum-exp over samples of the log-probs
                                   normal: sum over samples of the predicted means          */
    double* lps_lse;    /* [N]     log-sum-exp over samples of log p(y_n | sample)          */
    double* lps_last;   /* [N]     log p(y_n | most recent sample)                          */
    double* acc_last;   /* [N]     most recent sample: 1/0 correct, or squared error        */
    double* rows;       /* [N][BNNP_EVAL_ROW]  scratch of bnnp_eval_finish                  */
    int64_t N;
    int32_t C;          /* classes (categorical) or output dimensions (normal)              */
    int32_t kind;       /* BNNP_EVAL_*                                                      */
} BnnpEvalState;

/* One batch of one sample: rows n0 .. n0+B-1 of the test set.
 *   acc_data  [B][C] fp32, row stride `stride` floats: preds.logits (normalised
 *             log-probabilities, exp_utils.py:282) or preds.mean (:285)
 *   lps       [B] fp32 log p(y | sample) per point (:283,:286), or null for categorical:
 *             then it is acc_data[b][labels[b]], which is what Categorical.log_prob returns
 *   labels    categorical: [B] int64 class indices;  normal: null
 *   targets   normal: [B][C] fp32 (row stride `stride_t`);  categorical: null
 *   sample_index  0 for the first sample (initialises the accumulators of these rows)
 * Replaces exp_utils.py:279-297 for that batch. */
int bnnp_eval_batch(const BnnpEvalState* st, const float* acc_data, int64_t stride, const float* lps,
                    const int64_t* labels, const float* targets, int64_t stride_t, int64_t n0, int32_t B,
                    int32_t sample_index, void* stream);

/* After the last batch of the last sample: exp_utils.py:301-322 on the device.
 *   labels / targets  the whole test set ([N] int64 / [N][C] fp32 contiguous)
 *   n_samples         E
 *   out               [BNNP_EV_OUT] doubles (device)
 *   probs_mean        null, or [N][C] doubles: ensemble_preds.probs (:324, calibration) */
int bnnp_eval_finish(const BnnpEvalState* st, const int64_t* labels, const float* targets, int32_t n_samples,
                     double* out, double* probs_mean, void* stream);

/* Text of the last error of the two calls above (thread-local). */
const char* bnnp_eval_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* BNNP_EVAL_H */
