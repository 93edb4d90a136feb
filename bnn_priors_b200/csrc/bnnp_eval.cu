// bnnp_eval.cu -- device-side bookkeeping of evaluate_model (bnn_priors/exp_utils.py:250-340)
// behind the C ABI of include/bnnp_eval.h.
//
// The reference copies every test batch's log-probabilities and logits to the CPU as
// float64 (a blocking copy per batch), keeps [E, N] and [E, N, C] tensors there and
// reduces them at the end.  Here the per-point accumulators live in HBM: one launch
// per batch folds the batch into them (running log-sum-exp over the samples), one
// launch pair at the end produces the four numbers the runners log, all in float64 and
// in a fixed order (bit-reproducible).  Small, latency-bound work: one warp per test
// point, lanes over the classes.

#include "bnnp.h"
#include "bnnp_eval.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

namespace {

thread_local char g_eval_err[256] = "";

constexpr int EV_THREADS = 256;
constexpr int EV_WARPS = EV_THREADS / 32;
constexpr int ROW = BNNP_EVAL_ROW;   // doubles per test point in BnnpEvalState.rows
constexpr int FIN_THREADS = 1024;

__device__ __forceinline__ double log_add_exp(double a, double b) {
    const double m = fmax(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + log(exp(a - m) + exp(b - m));
}

// (value, index) argmax over the warp: larger value wins, ties go to the smaller index
__device__ __forceinline__ void warp_argmax(double& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// exp_utils.py:279-297 for one batch of one sample
__global__ void __launch_bounds__(EV_THREADS) eval_batch_kernel(const BnnpEvalState st, const float* __restrict__ acc,
                                                                int64_t stride, const float* __restrict__ lps,
                                                                const int64_t* __restrict__ labels,
                                                                const float* __restrict__ targets, int64_t stride_t,
                                                                int64_t n0, int B, int first) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * EV_WARPS + (threadIdx.x >> 5);
    if (b >= B) return;
    const int64_t n = n0 + b;
    const int C = st.C;
    const float* row = acc + (int64_t)b * stride;
    double* ens = st.ens + n * C;
    if (st.kind == BNNP_EVAL_CATEGORICAL) {
        double best = -INFINITY;
        int besti = INT32_MAX;
        for (int c = lane; c < C; c += 32) {
            const double x = (double)row[c];
            ens[c] = first ? x : log_add_exp(ens[c], x);
            if (x > best || (x == best && c < besti)) { best = x; besti = c; }
        }
        warp_argmax(best, besti);
        if (lane == 0) {
            const int64_t y = labels[b];
            const double lp = lps != nullptr ? (double)lps[b] : ((y >= 0 && y < C) ? (double)row[y] : NAN);
            st.acc_last[n] = (besti == (int)y) ? 1.0 : 0.0;      // models/base.py:184-185
            st.lps_last[n] = lp;
            st.lps_lse[n] = first ? lp : log_add_exp(st.lps_lse[n], lp);
        }
    } else {
        const float* t = targets + (int64_t)b * stride_t;
        double se = 0.0;
        for (int c = lane; c < C; c += 32) {
            const double x = (double)row[c];
            ens[c] = first ? x : ens[c] + x;
            const double d = x - (double)t[c];
            se += d * d;
        }
        se = warp_sum(se);
        if (lane == 0) {
            const double lp = (double)lps[b];
            st.acc_last[n] = se;                                  // models/base.py:155-158
            st.lps_last[n] = lp;
            st.lps_lse[n] = first ? lp : log_add_exp(st.lps_lse[n], lp);
        }
    }
}

// exp_utils.py:301-321 per test point
__global__ void __launch_bounds__(EV_THREADS) eval_rows_kernel(const BnnpEvalState st, const int64_t* __restrict__ labels,
                                                               const float* __restrict__ targets, int n_samples,
                                                               double* __restrict__ probs_mean) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * EV_WARPS + (threadIdx.x >> 5);
    if (n >= st.N) return;
    const int C = st.C;
    const double log_e = log((double)n_samples);
    const double* ens = st.ens + n * C;
    double acc_ens, lp_check = 0.0;
    if (st.kind == BNNP_EVAL_CATEGORICAL) {
        // _log_space_mean(acc_data, 0), then Categorical(logits=...) normalises it
        double best = -INFINITY;
        int besti = INT32_MAX;
        for (int c = lane; c < C; c += 32) {
            const double a = ens[c] - log_e;
            if (a > best || (a == best && c < besti)) { best = a; besti = c; }
        }
        warp_argmax(best, besti);
        double s = 0.0;
        for (int c = lane; c < C; c += 32) s += exp((ens[c] - log_e) - best);
        s = warp_sum(s);
        const double lse = best + log(s);
        const int64_t y = labels[n];
        if (probs_mean != nullptr)
            for (int c = lane; c < C; c += 32) probs_mean[n * C + c] = exp((ens[c] - log_e) - lse);
        acc_ens = (besti == (int)y) ? 1.0 : 0.0;
        lp_check = (y >= 0 && y < C) ? (ens[y] - log_e) - lse : NAN;
    } else {
        // Normal(acc_data.mean(0), 1): squared error of the ensemble mean
        const float* t = targets + n * C;
        double se = 0.0;
        for (int c = lane; c < C; c += 32) {
            const double d = ens[c] / (double)n_samples - (double)t[c];
            se += d * d;
        }
        acc_ens = warp_sum(se);
    }
    if (lane == 0) {
        double* r = st.rows + n * ROW;
        r[0] = st.lps_lse[n] - log_e;      // _log_space_mean(lps, 0)
        r[1] = st.lps_last[n];
        r[2] = acc_ens;
        r[3] = st.acc_last[n];
        r[4] = lp_check;
    }
}

// means over the test set, one CTA, fixed summation order
__global__ void __launch_bounds__(FIN_THREADS) eval_mean_kernel(const BnnpEvalState st, double* __restrict__ out) {
    __shared__ double sh[ROW][FIN_THREADS / 32];
    double s[ROW];
#pragma unroll
    for (int k = 0; k < ROW; ++k) s[k] = 0.0;
    for (int64_t n = threadIdx.x; n < st.N; n += FIN_THREADS) {
#pragma unroll
        for (int k = 0; k < ROW; ++k) s[k] += st.rows[n * ROW + k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ROW; ++k) {
        s[k] = warp_sum(s[k]);
        if (lane == 0) sh[k][warp] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < ROW) {
        double t = 0.0;
        for (int w = 0; w < FIN_THREADS / 32; ++w) t += sh[threadIdx.x][w];
        out[threadIdx.x] = t / (double)st.N;
    }
    if (threadIdx.x >= ROW && threadIdx.x < BNNP_EV_OUT) out[threadIdx.x] = 0.0;
}

int ev_fail(int code, const char* what) {
    snprintf(g_eval_err, sizeof(g_eval_err), "%s", what);
    return code;
}

int check_state(const BnnpEvalState* st, const char* who) {
    if (st == nullptr) return ev_fail(BNNP_E_ARG, who);
    if (st->ens == nullptr || st->lps_lse == nullptr || st->lps_last == nullptr || st->acc_last == nullptr ||
        st->rows == nullptr || st->N <= 0 || st->C <= 0)
        return ev_fail(BNNP_E_ARG, who);
    if (st->kind != BNNP_EVAL_CATEGORICAL && st->kind != BNNP_EVAL_NORMAL) return ev_fail(BNNP_E_ARG, who);
    return 0;
}

}  // namespace

extern "C" {

const char* bnnp_eval_last_error(void) { return g_eval_err; }

int bnnp_eval_batch(const BnnpEvalState* st, const float* acc_data, int64_t stride, const float* lps,
                    const int64_t* labels, const float* targets, int64_t stride_t, int64_t n0, int32_t B,
                    int32_t sample_index, void* stream) {
    if (int rc = check_state(st, "bnnp_eval_batch: bad state")) return rc;
    if (B == 0) return 0;
    if (acc_data == nullptr || B < 0 || n0 < 0 || n0 + B > st->N || stride < st->C || sample_index < 0)
        return ev_fail(BNNP_E_ARG, "bnnp_eval_batch: bad batch (rows outside the test set, or null data)");
    if (st->kind == BNNP_EVAL_CATEGORICAL && labels == nullptr)
        return ev_fail(BNNP_E_ARG, "bnnp_eval_batch: categorical predictions need labels");
    if (st->kind == BNNP_EVAL_NORMAL && (targets == nullptr || lps == nullptr || stride_t < st->C))
        return ev_fail(BNNP_E_ARG, "bnnp_eval_batch: normal predictions need targets and lps");
    const int blocks = (B + EV_WARPS - 1) / EV_WARPS;
    eval_batch_kernel<<<blocks, EV_THREADS, 0, (cudaStream_t)stream>>>(*st, acc_data, stride, lps, labels, targets,
                                                                       stride_t, n0, B, sample_index == 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_eval_err, sizeof(g_eval_err), "eval_batch_kernel launch: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int bnnp_eval_finish(const BnnpEvalState* st, const int64_t* labels, const float* targets, int32_t n_samples,
                     double* out, double* probs_mean, void* stream) {
    if (int rc = check_state(st, "bnnp_eval_finish: bad state")) return rc;
    if (out == nullptr || n_samples <= 0) return ev_fail(BNNP_E_ARG, "bnnp_eval_finish: bad argument");
    if (st->kind == BNNP_EVAL_CATEGORICAL && labels == nullptr)
        return ev_fail(BNNP_E_ARG, "bnnp_eval_finish: categorical predictions need labels");
    if (st->kind == BNNP_EVAL_NORMAL && targets == nullptr)
        return ev_fail(BNNP_E_ARG, "bnnp_eval_finish: normal predictions need targets");
    const int64_t blocks = (st->N + EV_WARPS - 1) / EV_WARPS;
    eval_rows_kernel<<<(unsigned)blocks, EV_THREADS, 0, (cudaStream_t)stream>>>(*st, labels, targets, n_samples,
                                                                                 probs_mean);
    eval_mean_kernel<<<1, FIN_THREADS, 0, (cudaStream_t)stream>>>(*st, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_eval_err, sizeof(g_eval_err), "eval finish launch: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

}  // extern "C"
