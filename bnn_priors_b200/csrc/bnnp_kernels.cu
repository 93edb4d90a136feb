// bnnp_kernels.cu -- the SG-MCMC sampler step for B200 (sm_100a) behind the C ABI
// declared in include/bnnp.h.
//
// One kernel family does every sampler transition of the reference
// (bnn_priors/mcmc/sgld.py:119-154, verlet_sgld.py:149-197, hmc.py:41-79,
// sample_momentum sgld.py:57-69) in ONE pass over the flat parameter / gradient /
// momentum arrays:  128-bit loads of p, g, m  ->  (optional) closed-form prior
// gradient (prior/loc_scale.py:34-77)  ->  noise from an in-register Philox4x32-10
// + Box-Muller (or a replay buffer for parity tests)  ->  momentum and parameter
// update  ->  128-bit stores of p', m' (and the verlet_sgld.py:72-83 snapshot)  ->
// eight dot products per chunk, reduced warp -> CTA in a fixed order and published
// as one fp64 partial record per chunk with plain stores (no fence, no atomics: a
// fence + ticket at the end of every CTA cost 9 us of an 88 us step,
// profiles/r01_tile_sweep.md).  The per-segment fold of those records and the
// sampler's scalar bookkeeping (delta_energy, prev_new_momentum_delta,
// est_temperature, est_config_temp, square_avg mean) are DEFERRED: the launch that
// follows carries the previous launch's epilogue parameters (BnnpLaunch.pending) and
// CTA j applies them for segment j before its own work; when the
// host wants a scalar, bnnp_finalize does the same in a tiny launch of its own.
// Either way it happens on the device, in a fixed order (bit-reproducible), and a
// step never synchronises the host.
//
// The path is HBM-bound elementwise work: no tensor cores, no shared-memory tiles;
// what matters is coalesced 16-byte accesses, enough loads in flight per SM, a lean
// instruction stream for the Philox rounds -- and, for chains larger than the 126 MB L2,
// what consecutive launches leave each other there: they walk the chain in opposite
// directions (BNNP_F_REVERSE) with evict-last P / M accesses.
//
// Also here: bnnp_prepass_kernel (the read-only pre-pass of the hierarchical priors,
// BNNP_F_HYPER), bnnp_finalize_kernel, bnnp_rollback_kernel and the host-side layout
// planner.  The evaluation kernels of include/bnnp_eval.h live in bnnp_eval.cu.

#include "bnnp.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace {

thread_local char g_err[256] = "";

int fail(int code, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

int fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

constexpr int THREADS = BNNP_THREADS;   // (-DBNNP_THREADS / -DBNNP_UNROLL override the header for tuning builds)
constexpr int UNROLL = BNNP_UNROLL;
#ifndef BNNP_MIN_CTAS
#define BNNP_MIN_CTAS 4   // resident CTAs per SM for the hot variants (measured: tools/tune_tiles.py,
#endif                    // profiles/r01_tile_sweep.md); 64 registers per thread at 256 threads
constexpr int CHUNK = BNNP_CHUNK;
constexpr int NWARPS = THREADS / 32;
static_assert(THREADS % 32 == 0 && THREADS >= 32, "whole warps");

// indices of the per-chunk partial sums
enum { R_GM_OLD = 0, R_GM_NEW, R_MM_OLD, R_MM_NEW, R_PG, R_GG, R_LOGP, R_NONFINITE };

// ---------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) and the Box-Muller transform specified in
// oracle/sgmcmc_oracle.py:_box_muller.  The round keys are uniform per launch and
// are precomputed on the host (StepParams).
// ---------------------------------------------------------------------------------
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};

// What the kernel receives: the caller's launch block plus values bnnp_launch derives from
// it on the host -- the ten Philox round keys and the fp32 roundings of the uniform
// coefficients.  They sit in the constant bank and are used as instruction operands
// directly (no registers, no per-round key arithmetic).
struct StepParams {
    BnnpLaunch L;
    PhiloxKeys keys;
    float cm, cn, gmax;
};

__host__ __device__ __forceinline__ PhiloxKeys philox_round_keys(uint32_t key0, uint32_t key1) {
    PhiloxKeys k;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = key0 + 0x9E3779B9u * (uint32_t)r;
        k.k1[r] = key1 + 0xBB67AE85u * (uint32_t)r;
    }
    return k;
}

__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                              const PhiloxKeys& k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // 32x32 -> 64-bit products: one IMAD.WIDE each
        const uint64_t p0 = (uint64_t)0xD2511F53u * (uint64_t)c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * (uint64_t)c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.k1[r];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#ifndef BNNP_LEAN_BM
#define BNNP_LEAN_BM 1
#endif
__device__ __forceinline__ float fast_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void box_muller(uint32_t x, uint32_t y, float& z0, float& z1) {
    const float u = fmaf(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0, 1]
    const float t = fmaf(__uint2float_rn(y), 2.3283064365386963e-10f, -0.5f);                     // [-.5, .5]
    const float theta = t * 6.283185307179586f;
    // sqrt(-2 ln u) = sqrt(-2 ln2 * log2 u); u >= 2^-33 is a normal number, so the bare MUFU.LG2 is enough
    // (what __logf does as well, minus its denormal / special-case handling)
#if BNNP_LEAN_BM
    const float r = fast_sqrt(-1.3862943611198906f * fast_lg2(u));
#else
    const float r = fast_sqrt(-2.0f * __logf(u));
#endif
    float s, c;
    __sincosf(theta, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// four N(0,1) for the flat element quad `quad` of launch `call`
__device__ __forceinline__ void philox_normal4(uint64_t quad, uint64_t call, const PhiloxKeys& k, float z[4]) {
    uint32_t c0 = (uint32_t)quad, c1 = (uint32_t)(quad >> 32);
    uint32_t c2 = (uint32_t)call, c3 = (uint32_t)(call >> 32);
    philox4x32_10(c0, c1, c2, c3, k);
    box_muller(c0, c1, z[0], z[1]);
    box_muller(c2, c3, z[2], z[3]);
}

// ---------------------------------------------------------------------------------
// Priors.  Per-segment constants (uniform over the CTA).
// ---------------------------------------------------------------------------------
// The kernel evaluates six closed forms; the ten prior kinds of the ABI map onto them:
//   F_NORMAL      Normal
//   F_LOGNORMAL   LogNormal (= Normal in p plus the "- p" of prior/loc_scale.py:91)
//   F_LAPLACE     Laplace
//   F_STUDENT_T   StudentT, Cauchy (= StudentT with df = 1)
//   F_GENNORM     GenNorm (prior/distributions.py:75-79)
//   F_DOUBLE_GAMMA DoubleGamma (prior/transformed.py:83-96)
//   F_NONE        no prior / Uniform / Improper: no gradient, a constant log density
//   F_CONST       hyper segments (one element u): the gradient term -(1/N) dlog p/du was left in
//                 the segment state by the epilogue of the BNNP_F_HYPER pre-pass
enum { F_NONE = 0, F_NORMAL, F_LOGNORMAL, F_LAPLACE, F_STUDENT_T, F_GENNORM, F_DOUBLE_GAMMA, F_CONST };

// Only the constants a form reads are set by make_prior<FORM>; the rest stay dead, so the
// common Normal case carries three floats (loc, k, inv_s) through the hot loop.
struct PriorConst {
    float loc;
    float k;        // F_NORMAL / F_LOGNORMAL: 1/(N s^2); F_LAPLACE: 1/(N s); others: 1/N
    float c;        // F_LOGNORMAL: the constant part of the gradient term, 1/N
    float a, b;     // F_STUDENT_T: a = df + 1, b = df s^2;  F_GENNORM: a = beta;  F_DOUBLE_GAMMA: a = conc - 1
    float la;       // F_STUDENT_T: -.5 (df + 1)
    float inv_s;    // 1/s
    float inv_df;
};

__device__ __forceinline__ int prior_form(int kind) {
    switch (kind) {
        case BNNP_PRIOR_NORMAL: return F_NORMAL;
        case BNNP_PRIOR_LOGNORMAL: return F_LOGNORMAL;
        case BNNP_PRIOR_LAPLACE: return F_LAPLACE;
        case BNNP_PRIOR_STUDENT_T:
        case BNNP_PRIOR_CAUCHY: return F_STUDENT_T;
        case BNNP_PRIOR_GENNORM: return F_GENNORM;
        case BNNP_PRIOR_DOUBLE_GAMMA: return F_DOUBLE_GAMMA;
        case BNNP_PRIOR_HYPER_GAMMA:
        case BNNP_PRIOR_HYPER_UNIFORM:
        case BNNP_PRIOR_HYPER_HALFCAUCHY:
        case BNNP_PRIOR_HYPER_IMPROPER: return F_CONST;
        default: return F_NONE;
    }
}

__host__ __device__ __forceinline__ bool is_hyper_kind(int kind) {
    return kind >= BNNP_PRIOR_HYPER_GAMMA && kind <= BNNP_PRIOR_HYPER_IMPROPER;
}
__host__ __device__ __forceinline__ bool may_have_hyper_scale(int kind) {
    return kind == BNNP_PRIOR_NORMAL || kind == BNNP_PRIOR_LAPLACE || kind == BNNP_PRIOR_STUDENT_T;
}

// Hierarchical priors: scale s(u) of a hyper segment and ds/du, float64.
//   softplus as torch.nn.functional.softplus (threshold 20): prior/transformed.py:57-58,74-75
//   a + b Phi(u): prior/transformed.py:28-30
__device__ __forceinline__ void hyper_scale(const BnnpSegment& h, double u, double& s, double& ds) {
    const double a = (double)h.prior_loc, b = (double)h.prior_scale;
    if (h.prior_kind == BNNP_PRIOR_HYPER_UNIFORM) {
        s = a + b * (0.5 * erfc(-u * 0.7071067811865476));
        ds = b * exp(-0.5 * u * u) * 0.3989422804014327;
        return;
    }
    const double sp = u > 20.0 ? u : log1p(exp(u));
    const double sg = u > 20.0 ? 1.0 : 1.0 / (1.0 + exp(-u));
    const double mult = h.prior_kind == BNNP_PRIOR_HYPER_HALFCAUCHY ? b : 1.0;
    s = sp * mult;
    ds = sg * mult;
}

// the same scale in float32 (what the reference's fp32 softplus / Normal.cdf give): used by every
// thread of a pre-pass CTA, where the float64 version above would dominate the launch
__device__ __forceinline__ float hyper_scale_f32(const BnnpSegment& h, float u) {
    if (h.prior_kind == BNNP_PRIOR_HYPER_UNIFORM)
        return fmaf(h.prior_scale, 0.5f * erfcf(-u * 0.70710678118654752f), h.prior_loc);
    const float sp = u > 20.0f ? u : log1pf(expf(u));
    return h.prior_kind == BNNP_PRIOR_HYPER_HALFCAUCHY ? sp * h.prior_scale : sp;
}

// `scale_free_log`: a NORMAL / LAPLACE segment with a sampled scale in a BNNP_F_HYPER_POST launch reduces
// sum(-d^2/2) resp. sum(-|d|) in the log-prior slot -- the statistic its hyper segment needs, independent
// of the scale this launch used -- instead of the log-density terms.
template <int FORM>
__device__ __forceinline__ PriorConst make_prior(const BnnpSegment& sd, double inv_n, float hyper_term,
                                                 bool scale_free_log = false) {
    PriorConst pc;
    pc.loc = sd.prior_loc;
    const float s = sd.prior_scale;
    pc.inv_s = scale_free_log ? 1.0f : 1.0f / s;
    pc.k = (float)inv_n;
    pc.c = pc.a = pc.b = pc.la = pc.inv_df = 0.0f;
    if (FORM == F_NORMAL || FORM == F_LOGNORMAL) {
        pc.k = (float)(inv_n / ((double)s * (double)s));
        pc.c = (float)inv_n;
    } else if (FORM == F_LAPLACE) {
        pc.k = (float)(inv_n / (double)s);
    } else if (FORM == F_STUDENT_T) {
        const float df = (sd.prior_kind == BNNP_PRIOR_CAUCHY) ? 1.0f : sd.prior_df;
        pc.inv_df = 1.0f / df;
        pc.a = df + 1.0f;
        pc.b = df * s * s;
        pc.la = -0.5f * (df + 1.0f);
    } else if (FORM == F_GENNORM) {
        pc.a = sd.prior_df;            // beta
    } else if (FORM == F_DOUBLE_GAMMA) {
        pc.a = sd.prior_df - 1.0f;     // concentration - 1
    } else if (FORM == F_CONST) {
        pc.c = hyper_term;
    }
    return pc;
}

// -(1/N) d log p / d theta: what potential = loss - log_prior/N (models/base.py:76)
// adds to p.grad through autograd in the reference.
template <int FORM>
__device__ __forceinline__ float prior_grad_term(const PriorConst& pc, float p) {
    const float d = p - pc.loc;
    if (FORM == F_NORMAL) return d * pc.k;
    if (FORM == F_LOGNORMAL) return fmaf(d, pc.k, pc.c);
    if (FORM == F_LAPLACE) return d == 0.0f ? 0.0f : copysignf(pc.k, d);
    if (FORM == F_STUDENT_T) return (pc.a * d) / fmaf(d, d, pc.b) * pc.k;
    if (FORM == F_GENNORM) {            // beta |z|^(beta-1) sign(d) / (s N)
        const float sg = d == 0.0f ? 0.0f : copysignf(1.0f, d);
        return pc.a * powf(fabsf(d) * pc.inv_s, pc.a - 1.0f) * sg * pc.inv_s * pc.k;
    }
    if (FORM == F_DOUBLE_GAMMA) {       // (sign(d)/s - (c-1)/d) / N
        const float sg = d == 0.0f ? 0.0f : copysignf(1.0f, d);
        return (sg * pc.inv_s - pc.a / d) * pc.k;
    }
    if (FORM == F_CONST) return pc.c;
    return 0.0f;
}

// the statistic of d log p / d scale a BNNP_F_HYPER launch reduces per linked segment
template <int FORM>
__device__ __forceinline__ float scale_stat_term(const PriorConst& pc, float p) {
    const float d = p - pc.loc;
    if (FORM == F_NORMAL) return d * d;
    if (FORM == F_LAPLACE) return fabsf(d);
    if (FORM == F_STUDENT_T) return (d * d) / fmaf(d, d, pc.b);
    return 0.0f;
}

// the theta-dependent part of log p(theta); the per-element constant is added once
// per segment in the epilogue (log_prior_const)
template <int FORM>
__device__ __forceinline__ float log_prior_term(const PriorConst& pc, float p) {
    const float d = p - pc.loc;
    const float z = d * pc.inv_s;
    if (FORM == F_NORMAL) return -0.5f * z * z;
    if (FORM == F_LOGNORMAL) return -0.5f * z * z - p;
    if (FORM == F_LAPLACE) return -fabsf(z);
    if (FORM == F_STUDENT_T) return pc.la * log1pf(z * z * pc.inv_df);
    if (FORM == F_GENNORM) return -powf(fabsf(z), pc.a);
    if (FORM == F_DOUBLE_GAMMA) return pc.a * logf(fabsf(d)) - fabsf(z);
    return 0.0f;
}

__device__ double log_prior_const(const BnnpSegment& sd) {
    const double s = (double)sd.prior_scale, df = (double)sd.prior_df;
    switch (sd.prior_kind) {
        case BNNP_PRIOR_NORMAL:
        case BNNP_PRIOR_LOGNORMAL: return -log(s) - 0.9189385332046727;   // .5 log 2pi
        case BNNP_PRIOR_LAPLACE: return -log(2.0 * s);
        case BNNP_PRIOR_STUDENT_T:
            return -(log(s) + 0.5 * log(df) + 0.5723649429247001 /* .5 log pi */ + lgamma(0.5 * df) -
                     lgamma(0.5 * (df + 1.0)));
        case BNNP_PRIOR_CAUCHY: return -log(s) - 1.1447298858494002;      // log pi
        case BNNP_PRIOR_GENNORM: return -log(2.0 * s) - lgamma(1.0 / df) + log(df);
        case BNNP_PRIOR_UNIFORM: return -log(s);                          // s = high - low
        case BNNP_PRIOR_DOUBLE_GAMMA: return -df * log(s) - lgamma(df) - 0.6931471805599453;
        default: return 0.0;                                              // NONE, IMPROPER
    }
}

// ---------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------
union F4 {
    float4 v;
    float f[4];
};

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// L2 eviction-priority hints (PTX createpolicy + ld/st .L2::cache_hint).  Consecutive
// launches walk the chain in opposite directions (BNNP_F_REVERSE), so what a launch leaves in the
// 126 MB L2 is what the next one reads first.  A parameter / momentum line kept there saves a DRAM
// read AND a write-back, so P and M are accessed evict-last; the snapshot stores (read again only
// after a rejection) are evict-first.  Measured (profiles/r01f_notes.md): serpentine order 79.1 ->
// 69.7 us per SGLD step, + evict-last P/M 67.5 us; evict-first gradient loads made no difference.
// The policy words come from createpolicy (ptxas folds the instruction with a constant fraction
// into an immediate, so this costs nothing over a hard-coded encoding and does not depend on one).
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#define POLICY_EVICT_FIRST policy_evict_first()
#define POLICY_EVICT_LAST policy_evict_last()
#ifndef BNNP_G_POLICY
#define BNNP_G_POLICY 0     // 0: plain loads, 1: evict-first
#endif
#ifndef BNNP_PM_POLICY
#define BNNP_PM_POLICY 2    // 0: plain, 1: evict-last stores, 2: evict-last loads and stores
#endif

__device__ __forceinline__ float4 ld_f4_hint(const float* p, uint64_t policy) {
    float4 v;
    asm("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void st_f4_hint(float* p, const float4& v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :
                 : "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ float4 ld_grad(const float* p) {
    return BNNP_G_POLICY == 1 ? ld_f4_hint(p, POLICY_EVICT_FIRST) : ld_f4(p);
}
// The quad that straddles the end of a tensor: only its valid floats are read.  (P, M and the flat G
// are padded to whole 128-byte lines, but a gradient tensor autograd handed over ends where it ends.)
__device__ __forceinline__ float4 ld_grad_tail(const float* p, int valid) {
    F4 v;
    v.v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (j < valid) v.f[j] = p[j];
    return v.v;
}
__device__ __forceinline__ float4 ld_state(const float* p) {
    return BNNP_PM_POLICY == 2 ? ld_f4_hint(p, POLICY_EVICT_LAST) : ld_f4(p);
}
__device__ __forceinline__ void st_state(float* p, const float4& v) {
    if (BNNP_PM_POLICY >= 1) st_f4_hint(p, v, POLICY_EVICT_LAST);
    else st_f4(p, v);
}

__device__ __forceinline__ double ld_cg_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Coef {
    float cm, cgM, cn, cpM, gmax;
};

// What changes from launch to launch.  Normally kernel parameters (constant bank); in capturable
// mode (BnnpLaunch.ctl) read from the chain's device control block, so that a captured launch can
// be replayed with a new Philox counter, the other walking direction and other coefficients.
struct Dyn {
    uint64_t call;
    int parity;
    uint32_t flags;
    float cm, cn;
    double cg, cp, inv_n;
};

__device__ __forceinline__ Dyn load_dyn(const StepParams& S) {
    const BnnpLaunch& L = S.L;
    Dyn d;
    if (L.ctl == nullptr) {
        d.call = L.call;
        d.parity = L.parity;
        d.flags = L.flags;
        d.cm = S.cm;
        d.cn = S.cn;
        d.cg = L.cg;
        d.cp = L.cp;
        d.inv_n = L.inv_num_data;
    } else {
        const BnnpControl* ctl = L.ctl;
        const BnnpCoef* cf = &ctl->coef[L.coef_slot];
        d.call = ctl->call;
        d.parity = ctl->parity;
        d.flags = d.parity ? (L.flags | BNNP_F_REVERSE) : (L.flags & ~(uint32_t)BNNP_F_REVERSE);
        d.cm = (float)cf->cm;
        d.cn = (float)cf->cn;
        d.cg = cf->cg;
        d.cp = cf->cp;
        d.inv_n = cf->inv_num_data;
    }
    return d;
}

// which dot products a launch needs (compile time: every one costs an FMA per element
// and ten shuffle steps per warp)
enum { SUMS_MIN = 0,      // g.g and the non-finite probe
       SUMS_VERLET = 1,   // + g.m, g.m'   (verlet_sgld.py:174-176)
       SUMS_ALL = 2 };    // + m.m, m'.m', p.g, sum log p

// One float4 of every stream.  Lanes >= `valid` (only the quad that straddles the end
// of a segment has valid < 4) arrive zeroed and stay zero.
template <int NOISE, bool PRIOR, int KIND, bool NOISE_FIRST, int SUMS, bool FULL>
__device__ __forceinline__ void update_quad(const uint32_t flags, const Coef& c, const PriorConst& pc, int valid,
                                            F4& p, const F4& g, F4& m, const float eps[4], float acc[BNNP_NRED]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float p0 = p.f[j], m0 = m.f[j];
        float gj = g.f[j];
        acc[R_NONFINITE] = fmaf(gj, 0.0f, acc[R_NONFINITE]);   // 0, or NaN once g is inf/NaN
        if (PRIOR) {
            if (KIND != F_NONE && (flags & BNNP_F_PRIOR_GRAD) && (FULL || j < valid)) gj += prior_grad_term<KIND>(pc, p0);
            if (flags & BNNP_F_CLAMP_GRAD) gj = fminf(fmaxf(gj, -c.gmax), c.gmax);
        }
        float t, pre;
        if (NOISE_FIRST) {           // verlet_sgld.py:163-167
            t = (NOISE != BNNP_NOISE_NONE) ? c.cn * eps[j] : 0.0f;
            t = fmaf(c.cgM, gj, t);
            if (c.cm != 0.0f) t = fmaf(c.cm, m0, t);
            pre = t;
        } else {                     // sgld.py:129,142 ; hmc.py:66 ; sgld.py:69
            t = c.cm * m0;
            t = fmaf(c.cgM, gj, t);
            pre = t;
            if (NOISE != BNNP_NOISE_NONE) t = fmaf(c.cn, eps[j], t);
        }
        acc[R_GG] = fmaf(gj, gj, acc[R_GG]);
        if (SUMS >= SUMS_VERLET) {
            acc[R_GM_OLD] = fmaf(gj, m0, acc[R_GM_OLD]);
            acc[R_GM_NEW] = fmaf(gj, t, acc[R_GM_NEW]);
        }
        if (SUMS >= SUMS_ALL) {
            const float mo = (flags & BNNP_F_MM_PRE_NOISE) ? pre : m0;
            acc[R_MM_OLD] = fmaf(mo, mo, acc[R_MM_OLD]);
            acc[R_MM_NEW] = fmaf(t, t, acc[R_MM_NEW]);
            acc[R_PG] = fmaf(p0, gj, acc[R_PG]);
        }
        const float pn = fmaf(c.cpM, t, p0);      // stored only with BNNP_F_WRITE_P
        if (PRIOR && KIND != F_NONE && KIND != F_CONST) {
            if ((flags & BNNP_F_LOG_PRIOR) && (FULL || j < valid))
                acc[R_LOGP] += log_prior_term<KIND>(pc, (flags & BNNP_F_WRITE_P) ? pn : p0);
        }
        p.f[j] = pn;
        m.f[j] = t;
    }
}

__host__ __device__ inline int sums_needed(int op, uint32_t flags) {
    if (flags & (BNNP_F_CALC_METRICS | BNNP_F_ALL_SUMS)) return SUMS_ALL;
    if (op == BNNP_OP_SAMPLE_MOMENTUM || op == BNNP_OP_REDUCE) return SUMS_ALL;
    return op == BNNP_OP_VERLET ? SUMS_VERLET : SUMS_MIN;
}

// What a hyper segment needs from its own value u and the statistic T of the segment it scales:
// the scale, its log density, -(1/N) d/du [ sum_i log p(w_i | s(u)) + log p_hyper(s(u)) ].
struct HyperValues {
    double s, lp, grad;
};
__device__ HyperValues hyper_values(const BnnpSegment& h, const BnnpSegment& w, double u, double T, double inv_num_data) {
    HyperValues v;
    double ds;
    hyper_scale(h, u, v.s, ds);
    const double s = v.s;
    const double a = (double)h.prior_loc, b = (double)h.prior_scale;
    double lp = 0.0, dlp = 0.0;
    if (h.prior_kind == BNNP_PRIOR_HYPER_GAMMA) {               // td.Gamma(a, b).log_prob(s)
        lp = a * log(b) + (a - 1.0) * log(s) - b * s - lgamma(a);
        dlp = ((a - 1.0) / s - b) * ds;
    } else if (h.prior_kind == BNNP_PRIOR_HYPER_UNIFORM) {      // -log(high - low)
        lp = -log(b);
    } else if (h.prior_kind == BNNP_PRIOR_HYPER_HALFCAUCHY) {   // td.HalfCauchy(a).log_prob(s)
        const double q = s / a;
        lp = -0.4515827052894548 /* log(2/pi) */ - log(a) - log1p(q * q);
        dlp = -(2.0 * q / a) / (1.0 + q * q) * ds;
    }
    const double n = (double)w.numel, df = (double)w.prior_df;
    double dl_ds = 0.0;                                          // sum_i d log p(w_i | s) / ds
    if (w.prior_kind == BNNP_PRIOR_NORMAL) dl_ds = T / (s * s * s) - n / s;
    else if (w.prior_kind == BNNP_PRIOR_LAPLACE) dl_ds = T / (s * s) - n / s;
    else if (w.prior_kind == BNNP_PRIOR_STUDENT_T) dl_ds = (df + 1.0) * T / s - n / s;
    v.lp = lp;
    v.grad = -inv_num_data * (dl_ds * ds + dlp);
    return v;
}

// BNNP_F_HYPER_POST: the value u' the launch `E` left in hyper segment `h`.  That launch's hyper CTA stashed
// it in the R_LOGP slot of its chunk's partial record (a hyper segment has no log-density term of its own
// there), where it stays put while the NEXT launch already rewrites P.
__device__ __forceinline__ double stashed_u(const BnnpLaunch& L, const BnnpEpilogue& E, const BnnpSegment& h) {
    return ld_cg_f64(L.partials + ((int64_t)E.parity * L.nchunks_total + h.first_chunk) * BNNP_NRED + R_LOGP);
}
// ... and the scale-free statistic of the segment it scales, from that launch's folded R_LOGP sum:
// the launch reduced sum(-d^2/2) (NORMAL) or sum(-|d|) (LAPLACE) for linked segments
__device__ __forceinline__ double post_statistic(const BnnpSegment& w, double folded) {
    return w.prior_kind == BNNP_PRIOR_NORMAL ? -2.0 * folded : -folded;
}

// Epilogue of the hierarchical-prior pre-pass (BNNP_F_HYPER) or of a BNNP_F_HYPER_POST step for one segment.
//   hyper segment h (link = weight segment w): from u and the statistic T of w (folded into r[BNNP_NRED]
//   by apply_pending):
//     BNNP_S_LOG_PRIOR <- log density of the scale           (scale_prior.log_prob())
//     BNNP_S_HYPER     <- -(1/N) d/du [ sum_i log p(w_i | s(u)) + log p_hyper(s(u)) ]
//     segs[w].prior_scale <- s(u)                            (what scale_prior() returns)
//   linked weight segment: its log-prior constant uses s(u); BNNP_S_HYPER <- T.
__device__ void hyper_epilogue(const BnnpLaunch& L, const BnnpEpilogue& E, const BnnpSegment& sd, double* st,
                               const double* r) {
    if (sd.link < 0 || sd.link >= L.nseg) return;
    // BNNP_F_HYPER_POST: the launch was a step; r[BNNP_NRED] is the linked segment's folded R_LOGP sum at the
    // parameters that launch left in P
    const bool post = (E.flags & BNNP_F_HYPER_POST) != 0;
    if (is_hyper_kind(sd.prior_kind)) {
        const BnnpSegment w = L.segs[sd.link];
        const double u = post ? stashed_u(L, E, sd) : (double)L.P[sd.off];
        const double T = post ? post_statistic(w, r[BNNP_NRED]) : r[BNNP_NRED];
        const HyperValues v = hyper_values(sd, w, u, T, E.inv_num_data);
        const double s = v.s, n = (double)w.numel;
        st[BNNP_S_LOG_PRIOR] = v.lp;
        st[BNNP_S_HYPER] = v.grad;
        if (post) {
            // the linked segment's log-prior at the new parameters AND the new scale
            double* st_w = L.seg_state + (int64_t)sd.link * BNNP_STATE_STRIDE;
            BnnpSegment now = w;
            now.prior_scale = (float)s;
            st_w[BNNP_S_LOG_PRIOR] = (w.prior_kind == BNNP_PRIOR_NORMAL ? -0.5 * T / (s * s) : -T / s) +
                                     n * log_prior_const(now);
            st_w[BNNP_S_HYPER] = T;
        }
        L.segs[sd.link].prior_scale = (float)s;
    } else if (may_have_hyper_scale(sd.prior_kind) && !post) {
        const BnnpSegment h = L.segs[sd.link];
        double s, ds;
        hyper_scale(h, (double)L.P[h.off], s, ds);
        BnnpSegment now = sd;
        now.prior_scale = (float)s;
        st[BNNP_S_LOG_PRIOR] = r[R_LOGP] + (double)sd.numel * log_prior_const(now);
        st[BNNP_S_HYPER] = r[R_GM_OLD];
    }
}

// The reference's per-tensor scalar bookkeeping for one segment and one launch `E`,
// applied to the segment-state array (fp64) from the eight folded sums `r`.
__device__ void segment_epilogue(const BnnpEpilogue& E, double* seg_state, const BnnpSegment& sd, int seg,
                                 const double* r) {
    const uint32_t flags = E.flags;
    const int sums = sums_needed(E.op, flags);
    const double gm_old = r[R_GM_OLD], gm_new = r[R_GM_NEW];
    const double mm_old = r[R_MM_OLD], mm_new = r[R_MM_NEW];
    const double pg = r[R_PG], gg = r[R_GG];
    double* st = seg_state + (int64_t)seg * BNNP_STATE_STRIDE;
    const double M = sd.precond;
    const bool metrics = (flags & BNNP_F_CALC_METRICS) != 0;

    if (sums >= SUMS_VERLET) {
        st[BNNP_S_GM_OLD] = gm_old;
        st[BNNP_S_GM_NEW] = gm_new;
    }
    if (sums == SUMS_ALL) {
        st[BNNP_S_MM_OLD] = mm_old;
        st[BNNP_S_MM_NEW] = mm_new;
    }
    if (flags & BNNP_F_READ_G) {
        st[BNNP_S_SUM_GG] = gg;
        st[BNNP_S_NONFINITE] = (r[R_NONFINITE] == 0.0) ? 0.0 : 1.0;
    }
    if (E.op == BNNP_OP_VERLET) {
        const double c_gm = E.c_gm_base * M;                       // verlet_sgld.py:170
        if (E.phase == BNNP_PHASE_INITIAL) {
            st[BNNP_S_DELTA_ENERGY] = -((M * M) * E.curv_base * gg);   // :171-172 with :44-47
        } else {
            double de = st[BNNP_S_DELTA_ENERGY];
            de += st[BNNP_S_PREV_NEW_MOM];                          // :174
            de += c_gm * gm_old;                                    // :175
            st[BNNP_S_DELTA_ENERGY] = de;
        }
        st[BNNP_S_PREV_NEW_MOM] = c_gm * gm_new;                    // :176
        if (metrics) st[BNNP_S_EST_MM] = (E.phase == BNNP_PHASE_FINAL) ? mm_new : mm_old;   // :181-187
    } else if (E.op == BNNP_OP_HMC) {
        if (E.phase == BNNP_PHASE_INITIAL && sums == SUMS_ALL) st[BNNP_S_DELTA_ENERGY] = -0.5 * mm_old;   // hmc.py:50-53
        if (metrics) st[BNNP_S_EST_MM] = (E.phase == BNNP_PHASE_FINAL) ? mm_new : mm_old;   // :55,60,72
    } else if (E.op == BNNP_OP_SGLD) {
        if (metrics) st[BNNP_S_EST_MM] = mm_old;                    // sgld.py:127,137
    }
    if (metrics && E.op <= BNNP_OP_HMC) st[BNNP_S_EST_PG] = pg;     // sgld.py:146
    if (sums == SUMS_ALL) {
        if (flags & BNNP_F_WRITE_M) st[BNNP_S_SUM_MM] = mm_new;
        else if (flags & BNNP_F_READ_M) st[BNNP_S_SUM_MM] = mm_old;
    }
    if (flags & BNNP_F_UPDATE_SQ)                                   // sgld.py:153-154, through its mean
        st[BNNP_S_SQ_MEAN] = E.rms_alpha * st[BNNP_S_SQ_MEAN] + (1.0 - E.rms_alpha) * (gg / (double)sd.numel);
    // (under BNNP_F_HYPER_POST the log-prior of a segment with a sampled scale is written by its hyper
    // segment's epilogue, at the new scale)
    if ((flags & BNNP_F_LOG_PRIOR) &&
        !((flags & BNNP_F_HYPER_POST) && sd.link >= 0 && may_have_hyper_scale(sd.prior_kind)))
        st[BNNP_S_LOG_PRIOR] = r[R_LOGP] + (double)sd.numel * log_prior_const(sd);
    st[BNNP_S_LAUNCHES] += 1.0;
}

// One warp: the sum over `num_chunks` partial records of one reduction slot (stride BNNP_NRED
// doubles).  Eight independent running sums per lane keep eight L2 loads in flight; they are
// combined in a fixed order, so the result does not depend on timing.
__device__ __forceinline__ double fold_records(const double* base, int num_chunks, int lane) {
    double s[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    for (int ch = lane; ch < num_chunks; ch += 32 * 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch + 32 * j;
            if (c < num_chunks) s[j] += ld_cg_f64(base + (int64_t)c * BNNP_NRED);
        }
    }
    const double t = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    return warp_sum_f64(t);
}

// the pending epilogue of a launch: from the kernel parameters, or from the control block
__device__ __forceinline__ BnnpEpilogue load_pending(const BnnpLaunch& L) {
    if (L.ctl == nullptr) return L.pending;
    return L.ctl->pending;
}

// CTA-wide: fold the partial records launch `E` left for segment `seg` (one warp
// per sum, lanes over the chunks, fp64, fixed order) and apply its epilogue.  A segment
// the pending launch skipped carries an older stamp and is left alone.
__device__ void apply_pending(const BnnpLaunch& L, const BnnpEpilogue& E, const BnnpSegment& sd, int seg,
                              double* s_sum) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t rec0 = (int64_t)E.parity * L.nchunks_total + sd.first_chunk;
    if (L.stamps[rec0] != E.call + 1) return;     // uniform over the CTA
    for (int k = warp; k < BNNP_NRED; k += NWARPS) {
        const double t = fold_records(L.partials + rec0 * BNNP_NRED + k, sd.num_chunks, lane);
        if (lane == 0) s_sum[k] = t;
    }
    if ((E.flags & (BNNP_F_HYPER | BNNP_F_HYPER_POST)) && is_hyper_kind(sd.prior_kind) && sd.link >= 0 &&
        sd.link < L.nseg && warp == 0) {
        // the statistic of the segment this hyper-parameter scales: fold that segment's records
        // here, so that no CTA depends on another CTA's epilogue
        const BnnpSegment w = L.segs[sd.link];
        const double* base = L.partials + ((int64_t)E.parity * L.nchunks_total + w.first_chunk) * BNNP_NRED +
                             ((E.flags & BNNP_F_HYPER_POST) ? R_LOGP : R_GM_OLD);
        const double t = fold_records(base, w.num_chunks, lane);
        if (lane == 0) s_sum[BNNP_NRED] = t;
    }
    __syncthreads();
    if (tid == 0) {
        segment_epilogue(E, L.seg_state, sd, seg, s_sum);
        if (E.flags & (BNNP_F_HYPER | BNNP_F_HYPER_POST)) hyper_epilogue(L, E, sd, L.seg_state + (int64_t)seg * BNNP_STATE_STRIDE, s_sum);
    }
    __syncthreads();   // s_sum is reused by the caller
}

// one 128-bit load of a BnnpChunk
struct ChunkDesc {
    int64_t fbase;
    int rem, seg;
};
static_assert(sizeof(BnnpChunk) == 16, "BnnpChunk is read with one 128-bit access");
__device__ __forceinline__ ChunkDesc load_chunk(const BnnpChunk* chunks, int chunk) {
    const int4 raw = *reinterpret_cast<const int4*>(chunks + chunk);
    ChunkDesc d;
    d.fbase = (int64_t)(((uint64_t)(uint32_t)raw.y << 32) | (uint64_t)(uint32_t)raw.x);
    d.rem = raw.z;
    d.seg = raw.w;
    return d;
}

struct ChunkCtx {
    int64_t fbase;   // flat index of the chunk's first float
    int rem;         // valid floats in this chunk
    int tid;
};

// Noise + update + stores for the UNROLL quads of one thread; the prior kind is a
// template parameter so the per-segment switch happens once per CTA.
template <int NOISE, bool PRIOR, int KIND, bool NOISE_FIRST, int SUMS, bool FULL>
__device__ __forceinline__ void process_chunk(const BnnpLaunch& L, const ChunkCtx& cx, const Coef& c,
                                              const PriorConst& pc, const PhiloxKeys& keys, F4 (&p)[UNROLL],
                                              F4 (&g)[UNROLL], F4 (&m)[UNROLL], F4 (&z)[UNROLL],
                                              float acc[BNNP_NRED], const uint32_t flags, const uint64_t call) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int e = (u * THREADS + cx.tid) * 4;
        if (!FULL && e >= cx.rem) continue;
        const int64_t fi = cx.fbase + e;
        if (flags & BNNP_F_SAVE_STATE) {   // verlet_sgld.py:72-83, the values BEFORE the update
            st_f4_hint(L.prev_p + fi, p[u].v, POLICY_EVICT_FIRST);
            st_f4_hint(L.prev_g + fi, g[u].v, POLICY_EVICT_FIRST);
            if (L.prev_m != nullptr) st_f4_hint(L.prev_m + fi, m[u].v, POLICY_EVICT_FIRST);
        }
        if (NOISE == BNNP_NOISE_PHILOX) philox_normal4((uint64_t)fi >> 2, call, keys, z[u].f);
        const int valid = FULL ? 4 : cx.rem - e;
        if (!FULL && valid < 4) {   // the quad that straddles the segment end: padding lanes are zeros
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (j >= valid) p[u].f[j] = g[u].f[j] = m[u].f[j] = z[u].f[j] = 0.0f;
        }
        update_quad<NOISE, PRIOR, KIND, NOISE_FIRST, SUMS, FULL>(flags, c, pc, valid, p[u], g[u], m[u], z[u].f, acc);
        if (flags & BNNP_F_WRITE_P) st_state(L.P + fi, p[u].v);
        if (flags & BNNP_F_WRITE_M) st_state(L.M + fi, m[u].v);
    }
}

// ---------------------------------------------------------------------------------
// The step kernel: one CTA per chunk of BNNP_CHUNK floats of one segment.
// ---------------------------------------------------------------------------------
// The production variants (no replay buffer, at most the Verlet sums) are held to
// BNNP_MIN_CTAS resident CTAs per SM; the metrics / replay variants need more
// registers and would only spill under that cap.
#ifndef BNNP_PRIOR_MIN_CTAS
#define BNNP_PRIOR_MIN_CTAS BNNP_MIN_CTAS
#endif

// ---------------------------------------------------------------------------------
// TMA staging (variants with more arithmetic per element).  A full chunk of a launch that reads all
// three streams is fetched by ONE thread with three 16 KB bulk copies (cp.async.bulk global -> shared,
// completion on an mbarrier) instead of twelve 128-bit loads per thread held in 48 registers: the
// threads then take one quad at a time out of shared memory, so the fused-prior and all-sums variants
// neither spill nor carry the 64-bit address arithmetic and predicates of the per-thread loads --
// they are issue-bound, not memory-bound (profiles/r02_ncu_variants.json).  Measured (profiles/r02_notes.md,
// production regime): metrics step 81.2 -> 78.5 us (back to 4 CTAs/SM), Verlet + fused prior 81.7 -> 80.9,
// Verlet + fused prior + metrics 91.0 -> 83.6; the lean variants gain nothing and keep the register path.
// ---------------------------------------------------------------------------------
#ifndef BNNP_TMA_MODE
#define BNNP_TMA_MODE 2     // 0: never; 1: every variant without a replay buffer; 2: fused-prior and all-sums variants
#endif
// What the threads can do without the data happens before they wait for it: the per-segment constants
// (make_prior: float64 divisions) and, BNNP_TMA_NOISE_AHEAD, the Philox + Box-Muller noise of that many of
// the thread's quads -- most of the arithmetic of a noisy step, now under the shadow of the bulk copies
// (profiles/r02z_tma_overlap.jsonl: Verlet + fused prior 72.1 -> 70.4 us back to back, + metrics 77.0 -> 73.8).
// BNNP_TMA_SUBTILES > 1 lets the chunk arrive as that many consecutive parts of every stream, each on an
// mbarrier of its own; measured slower (2: +0.5 us, 4: +1 .. +3 us), kept as an option at 1.
#ifndef BNNP_TMA_SUBTILES
#define BNNP_TMA_SUBTILES 1
#endif
#ifndef BNNP_TMA_NOISE_AHEAD
#define BNNP_TMA_NOISE_AHEAD 4
#endif
constexpr int TMA_SUBTILES = BNNP_TMA_SUBTILES;
static_assert(TMA_SUBTILES >= 1 && UNROLL % TMA_SUBTILES == 0, "sub-tiles are whole passes of the CTA over the chunk");
constexpr int TMA_STREAM_BYTES = CHUNK * 4;
constexpr int TMA_SUB_BYTES = TMA_STREAM_BYTES / TMA_SUBTILES;
constexpr int TMA_SMEM_BYTES = 3 * TMA_STREAM_BYTES + 8 * (TMA_SUBTILES < 2 ? 2 : TMA_SUBTILES);   // three staged streams + the mbarriers

template <int NOISE, bool PRIOR, int SUMS>
__host__ __device__ constexpr bool use_tma() {
    return NOISE != BNNP_NOISE_REPLAY &&
           (BNNP_TMA_MODE == 1 || (BNNP_TMA_MODE == 2 && (PRIOR || SUMS == 2)));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n"
                     "}"
                     : "=r"(done)
                     : "r"(mbar), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar), "l"(pol)
                 : "memory");
}

// process_chunk for a chunk staged in shared memory: one quad of every stream at a time
template <int NOISE, bool PRIOR, int KIND, bool NOISE_FIRST, int SUMS>
__device__ __forceinline__ void process_staged(const BnnpLaunch& L, const ChunkCtx& cx, const Coef& c, const PriorConst& pc,
                                               const PhiloxKeys& keys, const float* sP, const float* sG, const float* sM,
                                               const uint32_t mbar, float acc[BNNP_NRED], const uint32_t flags,
                                               const uint64_t call) {
    // The noise of the first BNNP_TMA_NOISE_AHEAD quads does not need the data: it is generated while the
    // bulk copies are in flight, which is most of the arithmetic of those quads.
    constexpr int AHEAD = (NOISE == BNNP_NOISE_PHILOX) ? (BNNP_TMA_NOISE_AHEAD < UNROLL ? BNNP_TMA_NOISE_AHEAD : UNROLL) : 0;
    F4 za[AHEAD > 0 ? AHEAD : 1];
#pragma unroll
    for (int u = 0; u < AHEAD; ++u)
        philox_normal4((uint64_t)(cx.fbase + (u * THREADS + cx.tid) * 4) >> 2, call, keys, za[u].f);
#pragma unroll
    for (int u = 0; u < AHEAD; ++u)      // pin the values in front of the wait (volatile asm statements keep their order)
#pragma unroll
        for (int j = 0; j < 4; ++j) asm volatile("" : "+f"(za[u].f[j]));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        constexpr int PER_SUB = UNROLL / TMA_SUBTILES;
        if (u % PER_SUB == 0) mbar_wait(mbar + 8 * (u / PER_SUB), 0);     // this pass's sub-tile has landed
        const int e = (u * THREADS + cx.tid) * 4;
        const int64_t fi = cx.fbase + e;
        F4 p, g, m, z;
        p.v = *reinterpret_cast<const float4*>(sP + e);
        g.v = *reinterpret_cast<const float4*>(sG + e);
        m.v = *reinterpret_cast<const float4*>(sM + e);
        z.v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (flags & BNNP_F_SAVE_STATE) {   // verlet_sgld.py:72-83, the values BEFORE the update
            st_f4_hint(L.prev_p + fi, p.v, POLICY_EVICT_FIRST);
            st_f4_hint(L.prev_g + fi, g.v, POLICY_EVICT_FIRST);
            if (L.prev_m != nullptr) st_f4_hint(L.prev_m + fi, m.v, POLICY_EVICT_FIRST);
        }
        if (u < AHEAD) z.v = za[u < AHEAD ? u : 0].v;
        else if (NOISE == BNNP_NOISE_PHILOX) philox_normal4((uint64_t)fi >> 2, call, keys, z.f);
        update_quad<NOISE, PRIOR, KIND, NOISE_FIRST, SUMS, true>(flags, c, pc, 4, p, g, m, z.f, acc);
        if (flags & BNNP_F_WRITE_P) st_state(L.P + fi, p.v);
        if (flags & BNNP_F_WRITE_M) st_state(L.M + fi, m.v);
    }
}

template <int NOISE, bool PRIOR, bool NOISE_FIRST, int SUMS>
__device__ __forceinline__ void staged_body(const BnnpLaunch& L, const PhiloxKeys& keys, const Dyn& dyn, const ChunkCtx& cx,
                                            const Coef& c, const BnnpSegment& sd, float hyper_term, bool scale_free,
                                            const float* sP, const float* sG, const float* sM, const uint32_t mbar,
                                            float acc[BNNP_NRED]) {
    const uint32_t flags = dyn.flags;
    if (PRIOR) {
#define BNNP_FORM_CASE(F)                                                                                       \
    case F:                                                                                                     \
        process_staged<NOISE, true, F, NOISE_FIRST, SUMS>(L, cx, c, make_prior<F>(sd, dyn.inv_n, hyper_term, scale_free), \
                                                          keys, sP, sG, sM, mbar, acc, flags, dyn.call);        \
        break;
        switch (prior_form(sd.prior_kind)) {
            BNNP_FORM_CASE(F_NORMAL)
#ifndef BNNP_ONLY_NORMAL
            BNNP_FORM_CASE(F_LOGNORMAL)
            BNNP_FORM_CASE(F_LAPLACE)
            BNNP_FORM_CASE(F_STUDENT_T)
            BNNP_FORM_CASE(F_GENNORM)
            BNNP_FORM_CASE(F_DOUBLE_GAMMA)
#endif
            BNNP_FORM_CASE(F_CONST)
            BNNP_FORM_CASE(F_NONE)
        }
#undef BNNP_FORM_CASE
    } else {
        process_staged<NOISE, false, F_NONE, NOISE_FIRST, SUMS>(L, cx, c, PriorConst(), keys, sP, sG, sM, mbar, acc, flags,
                                                                dyn.call);
    }
}

#ifndef BNNP_TMA_ALLSUMS_CTAS
#define BNNP_TMA_ALLSUMS_CTAS 4     // the all-sums variants fit 64 registers once the streams are staged in shared memory
#endif
template <int NOISE, int SUMS, bool PRIOR>
constexpr int min_ctas() {
    return (NOISE != BNNP_NOISE_REPLAY && SUMS != 2) ? (PRIOR ? BNNP_PRIOR_MIN_CTAS : BNNP_MIN_CTAS)
           : (use_tma<NOISE, PRIOR, SUMS>()          ? BNNP_TMA_ALLSUMS_CTAS
                                                     : (BNNP_MIN_CTAS > 3 ? 3 : BNNP_MIN_CTAS));
}

#ifndef BNNP_FULL_MODE
#define BNNP_FULL_MODE 2    // 0: general path only; 1: full-chunk fast path everywhere; 2: for SUMS_ALL variants
#endif
// Loads, noise, prior, update and stores of one chunk; the partial sums come back in `acc`.
template <int NOISE, bool PRIOR, bool NOISE_FIRST, int SUMS, bool FULL>
__device__ __forceinline__ void chunk_body(const BnnpLaunch& L, const PhiloxKeys& keys, const Dyn& dyn, const ChunkCtx& cx,
                                           const Coef& c, const BnnpSegment& sd, float hyper_term, bool scale_free,
                                           const float* gsrc, float acc[BNNP_NRED]) {
    const uint32_t flags = dyn.flags;
    const int tid = cx.tid;
    // ---- front-batched 128-bit loads: 3 (4 with replay noise) x UNROLL in flight per thread
    F4 p[UNROLL], g[UNROLL], m[UNROLL], z[UNROLL];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int e = (u * THREADS + tid) * 4;
        if (FULL) {
            p[u].v = ld_state(L.P + cx.fbase + e);
            g[u].v = ld_grad(gsrc + e);
            m[u].v = ld_state(L.M + cx.fbase + e);
            if (NOISE == BNNP_NOISE_REPLAY) z[u].v = ld_f4(L.replay_noise + cx.fbase + e);
            else z[u].v = zero4;
        } else {
            const bool act = e < cx.rem;
            p[u].v = (act && (flags & BNNP_F_READ_P)) ? ld_state(L.P + cx.fbase + e) : zero4;
            // whole quads only: P, M and the flat G are padded to whole 128-byte lines, but a gradient tensor
            // autograd handed over ends where it ends (the quad that straddles its end is fetched below)
            g[u].v = (e + 4 <= cx.rem && (flags & BNNP_F_READ_G)) ? ld_grad(gsrc + e) : zero4;
            m[u].v = (act && (flags & BNNP_F_READ_M)) ? ld_state(L.M + cx.fbase + e) : zero4;
            if (NOISE == BNNP_NOISE_REPLAY) z[u].v = act ? ld_f4(L.replay_noise + cx.fbase + e) : zero4;
            else z[u].v = zero4;
        }
    }
    if (!FULL && (cx.rem & 3) && (flags & BNNP_F_READ_G)) {
        // (uniform over the CTA: only the last chunk of a tensor whose size is not a multiple of 4)
        const int et = cx.rem & ~3;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if ((u * THREADS + tid) * 4 == et) g[u].v = ld_grad_tail(gsrc + et, cx.rem - et);
    }

    if (PRIOR) {
#define BNNP_FORM_CASE(F)                                                                                      \
    case F:                                                                                                    \
        process_chunk<NOISE, true, F, NOISE_FIRST, SUMS, FULL>(L, cx, c, make_prior<F>(sd, dyn.inv_n, hyper_term, scale_free), \
                                                               keys, p, g, m, z, acc, flags, dyn.call);        \
        break;
        switch (prior_form(sd.prior_kind)) {   // uniform over the CTA: one closed form per segment
            BNNP_FORM_CASE(F_NORMAL)
#ifndef BNNP_ONLY_NORMAL
            BNNP_FORM_CASE(F_LOGNORMAL)
            BNNP_FORM_CASE(F_LAPLACE)
            BNNP_FORM_CASE(F_STUDENT_T)
            BNNP_FORM_CASE(F_GENNORM)
            BNNP_FORM_CASE(F_DOUBLE_GAMMA)
#endif
            BNNP_FORM_CASE(F_CONST)
            BNNP_FORM_CASE(F_NONE)
        }
#undef BNNP_FORM_CASE
    } else {
        process_chunk<NOISE, false, F_NONE, NOISE_FIRST, SUMS, FULL>(L, cx, c, PriorConst(), keys, p, g, m, z, acc, flags,
                                                                     dyn.call);
    }
}

template <int NOISE, bool PRIOR, bool NOISE_FIRST, int SUMS>
__global__ void __launch_bounds__(THREADS, min_ctas<NOISE, SUMS, PRIOR>()) bnnp_step_kernel(const __grid_constant__ StepParams S) {
    __shared__ double s_red[NWARPS][BNNP_NRED];
    const BnnpLaunch& L = S.L;
    const PhiloxKeys& keys = S.keys;

    const int tid = threadIdx.x;
    const Dyn dyn = load_dyn(S);
    const uint32_t flags = dyn.flags;
    const int slot = (flags & BNNP_F_REVERSE) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    const int chunk = L.chunk_ids != nullptr ? L.chunk_ids[slot] : slot;
    const ChunkDesc cd = load_chunk(L.chunks, chunk);
    const int seg = cd.seg;
    constexpr uint32_t READ_ALL = BNNP_F_READ_P | BNNP_F_READ_G | BNNP_F_READ_M;
    constexpr bool TMA = use_tma<NOISE, PRIOR, SUMS>();
    const bool full = cd.rem == CHUNK && (flags & READ_ALL) == READ_ALL;
    extern __shared__ __align__(128) unsigned char bnnp_dsm[];
    float* const sP = reinterpret_cast<float*>(bnnp_dsm);
    float* const sG = sP + CHUNK;
    float* const sM = sG + CHUNK;
    const uint32_t mbar = smem_u32(sM + CHUNK);
    // thread t fetches sub-tile t (floats [t, t+1) * CHUNK / TMA_SUBTILES of every stream) onto mbarrier t
    const bool fetcher = TMA && full && tid < TMA_SUBTILES;
    constexpr int SUB = CHUNK / TMA_SUBTILES;
    if (fetcher) {
        // P and M do not wait for the segment descriptor: their bulk copies go out first
        mbar_init(mbar + 8 * tid, 1);
        mbar_arrive_expect_tx(mbar + 8 * tid, 3 * TMA_SUB_BYTES);
        bulk_g2s_hint(smem_u32(sP + tid * SUB), L.P + cd.fbase + tid * SUB, TMA_SUB_BYTES, mbar + 8 * tid, POLICY_EVICT_LAST);
        bulk_g2s_hint(smem_u32(sM + tid * SUB), L.M + cd.fbase + tid * SUB, TMA_SUB_BYTES, mbar + 8 * tid, POLICY_EVICT_LAST);
    }
    BnnpSegment sd = L.segs[seg];
    // the segment's gradient: its slice of the flat G, or the tensor autograd handed over
    const float* gsrc = L.seg_grad != nullptr ? L.seg_grad[seg] + (cd.fbase - sd.off) : L.G + cd.fbase;
    if (fetcher) bulk_g2s(smem_u32(sG + tid * SUB), gsrc + tid * SUB, TMA_SUB_BYTES, mbar + 8 * tid);

    // Sampled scales (hierarchical priors).  Normally the table holds the current scale of a linked segment
    // and the segment state the hyper-parameter's gradient term (left by the epilogue of a pre-pass or of
    // a BNNP_F_HYPER_POST step, finalised before this launch).  BNNP_F_HYPER_CHAIN: that epilogue is still
    // PENDING -- it rides on this launch like any other epilogue -- so the CTAs that need its results
    // derive them themselves from the pending launch's records: a linked segment its scale from the stashed
    // hyper-parameter, a hyper segment its gradient term from the folded statistic.  The CTA that applies
    // the epilogue (blockIdx = segment) writes the same values for the host; nobody waits for anybody.
    float hyper_term = 0.0f;
    bool scale_free = false;
    if (PRIOR) {
        __shared__ double s_chain[2];
        const bool linked = sd.link >= 0 && sd.link < L.nseg;
        const bool chain = (flags & BNNP_F_HYPER_CHAIN) != 0;
        scale_free = (flags & BNNP_F_HYPER_POST) && linked && may_have_hyper_scale(sd.prior_kind) &&
                     sd.prior_kind != BNNP_PRIOR_STUDENT_T;
        if (chain && linked && may_have_hyper_scale(sd.prior_kind)) {
            if (tid == 0) {
                const BnnpEpilogue E = load_pending(L);
                const BnnpSegment h = L.segs[sd.link];
                double sc, ds;
                hyper_scale(h, stashed_u(L, E, h), sc, ds);
                s_chain[0] = sc;
            }
            __syncthreads();
            sd.prior_scale = (float)s_chain[0];
        } else if (is_hyper_kind(sd.prior_kind)) {
            if (chain && linked) {
                const BnnpEpilogue E = load_pending(L);
                const BnnpSegment w = L.segs[sd.link];
                if (tid < 32) {
                    const double t = fold_records(L.partials + ((int64_t)E.parity * L.nchunks_total + w.first_chunk) * BNNP_NRED + R_LOGP,
                                                  w.num_chunks, tid);
                    if (tid == 0) s_chain[1] = hyper_values(sd, w, stashed_u(L, E, sd), post_statistic(w, t), E.inv_num_data).grad;
                }
                __syncthreads();
                hyper_term = (float)s_chain[1];
            } else {
                // -(1/N) d log p / du, left in the segment state by an epilogue that has been applied
                hyper_term = (float)L.seg_state[(int64_t)seg * BNNP_STATE_STRIDE + BNNP_S_HYPER];
            }
        }
    }
    // the previous launch's bookkeeping: segment j is handled by CTA j, i.e. by the CTAs that
    // start first, so the few microseconds it takes are absorbed at the front of the launch
    if ((int)blockIdx.x < L.nseg) {
        const BnnpEpilogue E = load_pending(L);
        if (E.valid) apply_pending(L, E, L.segs[blockIdx.x], blockIdx.x, s_red[0]);
    }
    ChunkCtx cx;
    cx.rem = cd.rem;
    cx.fbase = cd.fbase;
    cx.tid = tid;

    Coef c;
    c.cm = dyn.cm;
    c.cn = dyn.cn;
    c.cgM = (float)(dyn.cg * sd.precond);
    c.cpM = (float)(dyn.cp * sd.precond);
    c.gmax = S.gmax;

    float acc[BNNP_NRED];
#pragma unroll
    for (int k = 0; k < BNNP_NRED; ++k) acc[k] = 0.0f;

    if (TMA && full) {
        __syncthreads();                 // the mbarriers the fetching threads initialised are visible to everybody
        staged_body<NOISE, PRIOR, NOISE_FIRST, SUMS>(L, keys, dyn, cx, c, sd, hyper_term, scale_free, sP, sG, sM, mbar, acc);
    } else {
        // Nearly every chunk is a full one of a launch that reads all three streams: that case can run
        // without the per-quad bounds / flag tests and without the zero fill (FULL), the rest (the last
        // chunk of a tensor, launches that skip a stream) takes the general path.
        // (measured per variant, profiles/r02_notes.md: the split pays for the variants that reduce every sum)
        constexpr bool SPLIT = !TMA && (BNNP_FULL_MODE == 1 || (BNNP_FULL_MODE == 2 && SUMS == SUMS_ALL));
        if (SPLIT && full)
            chunk_body<NOISE, PRIOR, NOISE_FIRST, SUMS, true>(L, keys, dyn, cx, c, sd, hyper_term, scale_free, gsrc, acc);
        else
            chunk_body<NOISE, PRIOR, NOISE_FIRST, SUMS, false>(L, keys, dyn, cx, c, sd, hyper_term, scale_free, gsrc, acc);
    }

    // ---- chunk reduction: fp32 butterfly inside the warp, fp64 across warps (fixed order)
    constexpr unsigned LIVE = (SUMS == SUMS_ALL) ? 0xffu
                              : (SUMS == SUMS_VERLET) ? ((1u << R_GG) | (1u << R_NONFINITE) | (1u << R_GM_OLD) | (1u << R_GM_NEW))
                                                      : ((1u << R_GG) | (1u << R_NONFINITE));
#pragma unroll
    for (int k = 0; k < BNNP_NRED; ++k) {
        if (!((LIVE | (PRIOR ? (1u << R_LOGP) : 0u)) >> k & 1u)) continue;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < BNNP_NRED; ++k) s_red[warp][k] = (double)acc[k];
    }
    __syncthreads();
    if (tid < BNNP_NRED) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) s += s_red[w][tid];
        // BNNP_F_HYPER_POST: a hyper segment stashes the value it leaves in P where the next launch (which
        // may already be rewriting P) and the epilogue find it (stashed_u); the __syncthreads above made
        // thread 0's store visible
        if (PRIOR && tid == R_LOGP && (flags & BNNP_F_HYPER_POST) && is_hyper_kind(sd.prior_kind)) s = (double)L.P[sd.off];
        L.partials[((int64_t)dyn.parity * L.nchunks_total + chunk) * BNNP_NRED + tid] = s;
    }
    if (tid == 0) L.stamps[(int64_t)dyn.parity * L.nchunks_total + chunk] = dyn.call + 1;   // "this launch wrote it"
}

// ---------------------------------------------------------------------------------
// The pre-pass of the hierarchical priors (BNNP_F_HYPER): reads P only and reduces, per
// chunk, sum log p(theta | s(u)) (slot R_LOGP) and the statistic of d log p / d scale (slot
// R_GM_OLD), with the scale of a linked segment taken from the hyper-parameter u now in P.
// Same chunking, partial records and deferred epilogue as the step kernel, but 16 KB per CTA
// instead of 80 KB: what bounds it is the per-CTA latency chain (chunk -> segment -> hyper
// segment -> u), so it is kept lean enough for BNNP_PREPASS_CTAS resident CTAs per SM.
// ---------------------------------------------------------------------------------
#ifndef BNNP_PREPASS_CTAS
#define BNNP_PREPASS_CTAS 8
#endif

template <int FORM>
__device__ __forceinline__ void prepass_quads(const PriorConst& pc, const ChunkCtx& cx, const F4 (&p)[UNROLL],
                                              float& logp, float& stat) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int e = (u * THREADS + cx.tid) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (e + j < cx.rem) {
                logp += log_prior_term<FORM>(pc, p[u].f[j]);
                stat += scale_stat_term<FORM>(pc, p[u].f[j]);
            }
        }
    }
}

__global__ void __launch_bounds__(THREADS, BNNP_PREPASS_CTAS) bnnp_prepass_kernel(const __grid_constant__ StepParams S) {
    __shared__ double s_red[NWARPS][BNNP_NRED];
    const BnnpLaunch& L = S.L;
    const int tid = threadIdx.x;
    const Dyn dyn = load_dyn(S);
    const int chunk = (dyn.flags & BNNP_F_REVERSE) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    const ChunkDesc cd = load_chunk(L.chunks, chunk);
    const int seg = cd.seg;
    BnnpSegment sd = L.segs[seg];
    if ((int)blockIdx.x < L.nseg) {
        const BnnpEpilogue E = load_pending(L);
        if (E.valid) apply_pending(L, E, L.segs[blockIdx.x], blockIdx.x, s_red[0]);
    }
    ChunkCtx cx;
    cx.rem = cd.rem;
    cx.fbase = cd.fbase;
    cx.tid = tid;

    F4 p[UNROLL];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int e = (u * THREADS + tid) * 4;
        p[u].v = e < cx.rem ? ld_f4(L.P + cx.fbase + e) : zero4;
    }
    if (sd.link >= 0 && sd.link < L.nseg && may_have_hyper_scale(sd.prior_kind)) {
        const BnnpSegment h = L.segs[sd.link];
        sd.prior_scale = hyper_scale_f32(h, L.P[h.off]);
    }
    float logp = 0.0f, stat = 0.0f;
    switch (prior_form(sd.prior_kind)) {
#define BNNP_PRE_CASE(F) \
    case F: prepass_quads<F>(make_prior<F>(sd, dyn.inv_n, 0.0f), cx, p, logp, stat); break;
        BNNP_PRE_CASE(F_NORMAL)
        BNNP_PRE_CASE(F_LOGNORMAL)
        BNNP_PRE_CASE(F_LAPLACE)
        BNNP_PRE_CASE(F_STUDENT_T)
        BNNP_PRE_CASE(F_GENNORM)
        BNNP_PRE_CASE(F_DOUBLE_GAMMA)
#undef BNNP_PRE_CASE
        default: break;     // no prior, constant densities, hyper segments: the epilogue adds their constants
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        logp += __shfl_xor_sync(0xffffffffu, logp, o);
        stat += __shfl_xor_sync(0xffffffffu, stat, o);
    }
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) {
        s_red[warp][0] = (double)logp;
        s_red[warp][1] = (double)stat;
    }
    __syncthreads();
    if (tid < BNNP_NRED) {
        double s = 0.0;
        if (tid == R_LOGP || tid == R_GM_OLD) {
            const int k = tid == R_LOGP ? 0 : 1;
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) s += s_red[w][k];
        }
        L.partials[((int64_t)dyn.parity * L.nchunks_total + chunk) * BNNP_NRED + tid] = s;
    }
    if (tid == 0) L.stamps[(int64_t)dyn.parity * L.nchunks_total + chunk] = dyn.call + 1;
}

// bnnp_finalize: the pending epilogue of every segment, nothing else
__global__ void __launch_bounds__(THREADS) bnnp_finalize_kernel(const BnnpLaunch L) {
    __shared__ double s_red[BNNP_NRED + 1];
    const int seg = blockIdx.x;
    const BnnpEpilogue E = load_pending(L);
    if (E.valid) apply_pending(L, E, L.segs[seg], seg, s_red);
}

// Capturable mode: the control block moves on after a launch (bnnp_advance).
__global__ void bnnp_advance_kernel(BnnpControl* ctl, int op, int phase, uint32_t flags, int coef_slot) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const BnnpCoef cf = ctl->coef[coef_slot];
    BnnpEpilogue E;
    E.valid = 1;
    E.op = op;
    E.phase = phase;
    E.parity = ctl->parity;
    E.flags = E.parity ? (flags | BNNP_F_REVERSE) : (flags & ~(uint32_t)BNNP_F_REVERSE);
    E.reserved = 0;
    E.call = ctl->call;
    E.c_gm_base = cf.c_gm_base;
    E.curv_base = cf.curv_base;
    E.rms_alpha = cf.rms_alpha;
    E.inv_num_data = cf.inv_num_data;
    ctl->pending = E;
    ctl->call = E.call + 1;
    ctl->parity = E.parity ^ 1;
}

__global__ void bnnp_clear_pending_kernel(BnnpControl* ctl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) ctl->pending.valid = 0;
}

// bnnp_poke: the payload travels in the kernel parameters
template <int WORDS>
struct PokePayload {
    uint32_t w[WORDS];
};
template <int WORDS>
__global__ void bnnp_poke_kernel(uint32_t* dst, const __grid_constant__ PokePayload<WORDS> src, int nwords) {
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src.w[i];
}

// P,G,M <- prev_* : verlet_sgld.py:63-69.  Three streams copied (24 B/param); UNROLL quads of every stream are
// loaded before the first store, like the step kernel's front-batched loads.
__global__ void __launch_bounds__(THREADS) bnnp_rollback_kernel(float* __restrict__ P, float* __restrict__ G,
                                                                float* __restrict__ M,
                                                                const float* __restrict__ pp,
                                                                const float* __restrict__ pg,
                                                                const float* __restrict__ pm, int64_t nquads) {
    const int64_t stride = (int64_t)gridDim.x * THREADS * UNROLL;
    for (int64_t q0 = (int64_t)blockIdx.x * THREADS * UNROLL + threadIdx.x; q0 < nquads; q0 += stride) {
        float4 a[UNROLL], b[UNROLL], cc[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t q = q0 + (int64_t)u * THREADS;
            if (q < nquads) {
                a[u] = ld_f4(pp + 4 * q);
                b[u] = ld_f4(pg + 4 * q);
                if (pm != nullptr) cc[u] = ld_f4(pm + 4 * q);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t q = q0 + (int64_t)u * THREADS;
            if (q < nquads) {
                st_f4(P + 4 * q, a[u]);
                st_f4(G + 4 * q, b[u]);
                if (pm != nullptr) st_f4(M + 4 * q, cc[u]);
            }
        }
    }
}

// Diagnostic: the bare access pattern of a step (read p, g, m; write p, m) with two FMAs
// per element and nothing else -- the ceiling tools/tune_tiles.py compares the real kernel to.
__global__ void __launch_bounds__(THREADS, BNNP_MIN_CTAS) bnnp_probe_stream_kernel(float* __restrict__ P,
                                                                                   const float* __restrict__ G,
                                                                                   float* __restrict__ M,
                                                                                   int64_t nquads) {
    const int64_t base = (int64_t)blockIdx.x * (THREADS * UNROLL) + threadIdx.x;
    float4 p[UNROLL], g[UNROLL], m[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int64_t q = base + (int64_t)u * THREADS;
        if (q < nquads) {
            p[u] = ld_f4(P + 4 * q);
            g[u] = ld_f4(G + 4 * q);
            m[u] = ld_f4(M + 4 * q);
        }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int64_t q = base + (int64_t)u * THREADS;
        if (q < nquads) {
            m[u].x = fmaf(0.5f, g[u].x, m[u].x); m[u].y = fmaf(0.5f, g[u].y, m[u].y);
            m[u].z = fmaf(0.5f, g[u].z, m[u].z); m[u].w = fmaf(0.5f, g[u].w, m[u].w);
            p[u].x = fmaf(0.5f, m[u].x, p[u].x); p[u].y = fmaf(0.5f, m[u].y, p[u].y);
            p[u].z = fmaf(0.5f, m[u].z, p[u].z); p[u].w = fmaf(0.5f, m[u].w, p[u].w);
            st_f4(P + 4 * q, p[u]);
            st_f4(M + 4 * q, m[u]);
        }
    }
}

typedef void (*StepKernel)(const StepParams);

template <int NOISE, bool PRIOR, bool NF>
StepKernel pick_sums(int sums) {
    switch (sums) {
        case SUMS_MIN: return bnnp_step_kernel<NOISE, PRIOR, NF, SUMS_MIN>;
        case SUMS_VERLET: return bnnp_step_kernel<NOISE, PRIOR, NF, SUMS_VERLET>;
        case SUMS_ALL: return bnnp_step_kernel<NOISE, PRIOR, NF, SUMS_ALL>;
    }
    return nullptr;
}

template <int NOISE>
StepKernel pick_noise(bool prior, bool noise_first, int sums) {
    if (prior) return noise_first ? pick_sums<NOISE, true, true>(sums) : pick_sums<NOISE, true, false>(sums);
    return noise_first ? pick_sums<NOISE, false, true>(sums) : pick_sums<NOISE, false, false>(sums);
}

}  // namespace

// The 36 instantiations of the step kernel take most of the build time, so the file can be compiled
// as three translation units in parallel (bnn_priors_b200/build.py): -DBNNP_PART=0 / 1 / 2 keeps the
// instantiations of one noise kind each (part 0 also holds the API); without BNNP_PART everything is
// in one unit.  The entry points below have external linkage and return the kernel as void*.
#ifndef BNNP_PART
#define BNNP_PART -1
#endif
void* bnnp_pick_noise_none(bool prior, bool noise_first, int sums);
void* bnnp_pick_noise_replay(bool prior, bool noise_first, int sums);
void* bnnp_pick_noise_philox(bool prior, bool noise_first, int sums);
#if BNNP_PART == -1 || BNNP_PART == 0
void* bnnp_pick_noise_none(bool prior, bool noise_first, int sums) {
    return (void*)pick_noise<BNNP_NOISE_NONE>(prior, noise_first, sums);
}
#endif
#if BNNP_PART == -1 || BNNP_PART == 1
void* bnnp_pick_noise_replay(bool prior, bool noise_first, int sums) {
    return (void*)pick_noise<BNNP_NOISE_REPLAY>(prior, noise_first, sums);
}
#endif
#if BNNP_PART == -1 || BNNP_PART == 2
void* bnnp_pick_noise_philox(bool prior, bool noise_first, int sums) {
    return (void*)pick_noise<BNNP_NOISE_PHILOX>(prior, noise_first, sums);
}
#endif

#if BNNP_PART == -1 || BNNP_PART == 0
namespace {

StepKernel pick_kernel(int noise, bool prior, bool noise_first, int sums) {
    switch (noise) {
        case BNNP_NOISE_NONE: return (StepKernel)bnnp_pick_noise_none(prior, noise_first, sums);
        case BNNP_NOISE_REPLAY: return (StepKernel)bnnp_pick_noise_replay(prior, noise_first, sums);
        case BNNP_NOISE_PHILOX: return (StepKernel)bnnp_pick_noise_philox(prior, noise_first, sums);
    }
    return nullptr;
}

bool misaligned(const void* p) { return ((uintptr_t)p & 15u) != 0; }

bool variant_uses_tma(int noise, bool prior, int sums) {
    return noise != BNNP_NOISE_REPLAY && (BNNP_TMA_MODE == 1 || (BNNP_TMA_MODE == 2 && (prior || sums == 2)));
}

// dynamic shared memory of a step-kernel instantiation (the TMA-staged variants need 48 KB + 16 B, above the
// default limit: the attribute is set once per instantiation and device)
int step_kernel_smem(StepKernel k, int noise, bool prior, int sums) {
    if (!variant_uses_tma(noise, prior, sums)) return 0;
    static thread_local const void* done[64];
    static thread_local int ndone = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const void* tag = (const void*)((uintptr_t)(const void*)k ^ ((uintptr_t)dev << 56));
    for (int i = 0; i < ndone; ++i)
        if (done[i] == tag) return TMA_SMEM_BYTES;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES) != cudaSuccess) return -1;
    if (ndone < 64) done[ndone++] = tag;
    return TMA_SMEM_BYTES;
}

}  // namespace

extern "C" {

int bnnp_abi_version(void) { return BNNP_ABI_VERSION; }

const char* bnnp_last_error(void) { return g_err; }

int bnnp_device_info(int device, int* sm_count, int* l2_bytes) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceProperties");
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = prop.l2CacheSize;
    return 0;
}

int bnnp_max_ctas_per_sm(int noise, int has_prior, int noise_first, int sums, int* out) {
    StepKernel k = pick_kernel(noise, has_prior != 0, noise_first != 0, sums);
    if (k == nullptr || out == nullptr) return fail(BNNP_E_ARG, "bnnp_max_ctas_per_sm: bad variant");
    const int smem = step_kernel_smem(k, noise, has_prior != 0, sums);
    if (smem < 0) return fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k, THREADS, smem);
    if (e != cudaSuccess) return fail_cuda(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    return 0;
}

int bnnp_plan_layout(const int64_t* numel, int nseg, int64_t* off, int32_t* first_chunk, int32_t* num_chunks,
                     int64_t* total_elems, int32_t* total_chunks, BnnpChunk* chunks) {
    if (numel == nullptr || nseg < 0 || off == nullptr || first_chunk == nullptr || num_chunks == nullptr)
        return fail(BNNP_E_ARG, "bnnp_plan_layout: null argument");
    int64_t o = 0, ch = 0;
    for (int i = 0; i < nseg; ++i) {
        if (numel[i] <= 0) return fail(BNNP_E_ARG, "bnnp_plan_layout: segments must not be empty");
        const int64_t nch = (numel[i] + BNNP_CHUNK - 1) / BNNP_CHUNK;
        if (ch + nch > INT32_MAX) return fail(BNNP_E_ARG, "bnnp_plan_layout: too many chunks");
        off[i] = o;
        first_chunk[i] = (int32_t)ch;
        num_chunks[i] = (int32_t)nch;
        if (chunks != nullptr)
            for (int64_t k = 0; k < nch; ++k) {
                const int64_t left = numel[i] - k * BNNP_CHUNK;
                chunks[ch + k].fbase = o + k * BNNP_CHUNK;
                chunks[ch + k].rem = left < (int64_t)BNNP_CHUNK ? (int32_t)left : BNNP_CHUNK;
                chunks[ch + k].seg = i;
            }
        ch += nch;
        o += (numel[i] + BNNP_SEG_ALIGN - 1) / BNNP_SEG_ALIGN * BNNP_SEG_ALIGN;
    }
    if (total_elems) *total_elems = o;
    if (total_chunks) *total_chunks = (int32_t)ch;
    return 0;
}

int bnnp_launch(const BnnpLaunch* a, void* stream) {
    if (a == nullptr) return fail(BNNP_E_ARG, "bnnp_launch: null args");
    if (a->nseg <= 0 || a->nchunks <= 0) return fail(BNNP_E_ARG, "bnnp_launch: empty chain");
    if (a->segs == nullptr || a->chunks == nullptr || a->seg_state == nullptr || a->partials == nullptr ||
        a->stamps == nullptr)
        return fail(BNNP_E_ARG, "bnnp_launch: null table pointer");
    if (a->nchunks > a->nchunks_total || (a->chunk_ids == nullptr && a->nchunks != a->nchunks_total))
        return fail(BNNP_E_ARG, "bnnp_launch: nchunks does not match the plan");
    // with a control block, parity / call / pending / coefficients live on the device (the caller keeps
    // the protocol: bnnp_advance after every launch, bnnp_finalize + bnnp_clear_pending before a launch
    // with chunk_ids or after a BNNP_F_HYPER / BNNP_F_HYPER_POST launch)
    const bool host_state = a->ctl == nullptr;
    if (!host_state && (a->coef_slot < 0 || a->coef_slot >= BNNP_COEF_SLOTS))
        return fail(BNNP_E_ARG, "bnnp_launch: coef_slot out of range");
    if (host_state && ((a->parity | 1) != 1 || (a->pending.valid && (a->pending.parity | 1) != 1)))
        return fail(BNNP_E_ARG, "bnnp_launch: parity must be 0 or 1");
    if (host_state && a->pending.valid && a->chunk_ids != nullptr)
        return fail(BNNP_E_ARG, "bnnp_launch: a launch with chunk_ids cannot carry a pending epilogue; bnnp_finalize first");
    if (host_state && a->pending.valid && a->pending.parity == a->parity)
        return fail(BNNP_E_ARG, "bnnp_launch: this launch would overwrite the partial records of the pending one");
    if (a->op < BNNP_OP_SGLD || a->op > BNNP_OP_REDUCE) return fail(BNNP_E_ARG, "bnnp_launch: bad op");
    if (a->phase < BNNP_PHASE_INITIAL || a->phase > BNNP_PHASE_FINAL) return fail(BNNP_E_ARG, "bnnp_launch: bad phase");
    const uint32_t f = a->flags;
    if ((f & (BNNP_F_READ_P | BNNP_F_WRITE_P)) && a->P == nullptr) return fail(BNNP_E_ARG, "bnnp_launch: P is null");
    if ((f & BNNP_F_READ_G) && a->G == nullptr && a->seg_grad == nullptr)
        return fail(BNNP_E_ARG, "bnnp_launch: G and seg_grad are both null");
    if ((f & (BNNP_F_READ_M | BNNP_F_WRITE_M)) && a->M == nullptr) return fail(BNNP_E_ARG, "bnnp_launch: M is null");
    if ((f & BNNP_F_WRITE_P) && !(f & BNNP_F_READ_P)) return fail(BNNP_E_ARG, "bnnp_launch: WRITE_P needs READ_P");
    if ((f & BNNP_F_SAVE_STATE) && (a->prev_p == nullptr || a->prev_g == nullptr))
        return fail(BNNP_E_ARG, "bnnp_launch: SAVE_STATE needs prev_p and prev_g");
    if (a->noise == BNNP_NOISE_REPLAY && a->replay_noise == nullptr)
        return fail(BNNP_E_ARG, "bnnp_launch: replay noise is null");
    if (misaligned(a->P) || misaligned(a->G) || misaligned(a->M) || misaligned(a->prev_p) || misaligned(a->prev_g) ||
        misaligned(a->prev_m) || misaligned(a->replay_noise) || misaligned(a->chunks))
        return fail(BNNP_E_ALIGN, "bnnp_launch: flat arrays and the chunk table must be 16-byte aligned");
    if (host_state && a->pending.valid && (a->pending.flags & (BNNP_F_HYPER | BNNP_F_HYPER_POST)) &&
        !((a->pending.flags & BNNP_F_HYPER_POST) && (f & BNNP_F_HYPER_CHAIN)))
        return fail(BNNP_E_ARG, "bnnp_launch: the epilogue of a BNNP_F_HYPER / BNNP_F_HYPER_POST launch rewrites the "
                                "segment table; bnnp_finalize first (or, for BNNP_F_HYPER_POST, launch with BNNP_F_HYPER_CHAIN)");
    if ((f & BNNP_F_HYPER_CHAIN) && host_state && !(a->pending.valid && (a->pending.flags & BNNP_F_HYPER_POST)))
        return fail(BNNP_E_ARG, "bnnp_launch: BNNP_F_HYPER_CHAIN needs a pending BNNP_F_HYPER_POST epilogue");
    if ((f & BNNP_F_HYPER_CHAIN) && ((f & BNNP_F_HYPER) || a->chunk_ids != nullptr))
        return fail(BNNP_E_ARG, "bnnp_launch: BNNP_F_HYPER_CHAIN is for launches over all chunks other than the pre-pass");
    if ((f & BNNP_F_HYPER_POST) &&
        ((f & BNNP_F_HYPER) || a->chunk_ids != nullptr ||
         (f & (BNNP_F_WRITE_P | BNNP_F_PRIOR_GRAD | BNNP_F_LOG_PRIOR)) !=
             (BNNP_F_WRITE_P | BNNP_F_PRIOR_GRAD | BNNP_F_LOG_PRIOR)))
        return fail(BNNP_E_ARG, "bnnp_launch: BNNP_F_HYPER_POST needs a step over all chunks with "
                                "WRITE_P | PRIOR_GRAD | LOG_PRIOR");
    if ((f & BNNP_F_HYPER) &&
        (a->op != BNNP_OP_REDUCE || !(f & BNNP_F_LOG_PRIOR) || !(f & BNNP_F_READ_P) || a->chunk_ids != nullptr ||
         (f & ~(uint32_t)(BNNP_F_HYPER | BNNP_F_LOG_PRIOR | BNNP_F_READ_P | BNNP_F_REVERSE))))
        return fail(BNNP_E_ARG, "bnnp_launch: BNNP_F_HYPER is the read-only pre-pass: BNNP_OP_REDUCE over all "
                                "chunks with exactly READ_P | LOG_PRIOR | HYPER");
    const bool prior = (f & (BNNP_F_LOG_PRIOR | BNNP_F_PRIOR_GRAD)) != 0;
    if (prior && !(f & BNNP_F_READ_P)) return fail(BNNP_E_ARG, "bnnp_launch: the prior needs READ_P");
    StepKernel k = (f & BNNP_F_HYPER) ? bnnp_prepass_kernel
                                      : pick_kernel(a->noise, prior, (f & BNNP_F_NOISE_FIRST) != 0, sums_needed(a->op, f));
    if (k == nullptr) return fail(BNNP_E_ARG, "bnnp_launch: bad noise kind");
    StepParams sp;
    sp.L = *a;
    sp.keys = philox_round_keys(a->key0, a->key1);
    sp.cm = (float)a->cm;
    sp.cn = (float)a->cn;
    sp.gmax = (float)a->grad_max;
    const int smem = (f & BNNP_F_HYPER) ? 0 : step_kernel_smem(k, a->noise, prior, sums_needed(a->op, f));
    if (smem < 0) return fail_cuda(cudaGetLastError(), "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    k<<<a->nchunks, THREADS, smem, (cudaStream_t)stream>>>(sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "bnnp_step_kernel launch");
    return 0;
}

int bnnp_finalize(const BnnpLaunch* a, void* stream) {
    if (a == nullptr) return fail(BNNP_E_ARG, "bnnp_finalize: null args");
    if (a->ctl == nullptr && !a->pending.valid) return 0;
    if (a->nseg <= 0 || a->segs == nullptr || a->seg_state == nullptr || a->partials == nullptr || a->stamps == nullptr)
        return fail(BNNP_E_ARG, "bnnp_finalize: null table pointer");
    if (a->ctl == nullptr && (a->pending.flags & (BNNP_F_HYPER | BNNP_F_HYPER_POST)) && a->P == nullptr)
        return fail(BNNP_E_ARG, "bnnp_finalize: the epilogue of a BNNP_F_HYPER / BNNP_F_HYPER_POST launch reads P");
    if (a->ctl == nullptr && (a->pending.parity | 1) != 1) return fail(BNNP_E_ARG, "bnnp_finalize: parity must be 0 or 1");
    bnnp_finalize_kernel<<<a->nseg, THREADS, 0, (cudaStream_t)stream>>>(*a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "bnnp_finalize_kernel launch");
    return 0;
}

int bnnp_advance(const BnnpLaunch* a, void* stream) {
    if (a == nullptr || a->ctl == nullptr) return fail(BNNP_E_ARG, "bnnp_advance: no control block");
    if (a->coef_slot < 0 || a->coef_slot >= BNNP_COEF_SLOTS) return fail(BNNP_E_ARG, "bnnp_advance: coef_slot out of range");
    bnnp_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a->ctl, a->op, a->phase, a->flags, a->coef_slot);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "bnnp_advance_kernel launch");
    return 0;
}

int bnnp_clear_pending(BnnpControl* ctl, void* stream) {
    if (ctl == nullptr) return fail(BNNP_E_ARG, "bnnp_clear_pending: no control block");
    bnnp_clear_pending_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ctl);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "bnnp_clear_pending_kernel launch");
    return 0;
}

int bnnp_poke(void* dst, const void* src_host, int64_t nbytes, void* stream) {
    if (dst == nullptr || src_host == nullptr || nbytes <= 0 || nbytes % 4 != 0 || nbytes > 3840 ||
        ((uintptr_t)dst & 3u) != 0)
        return fail(BNNP_E_ARG, "bnnp_poke: need 0 < nbytes <= 3840, a multiple of 4, and a 4-byte aligned destination");
    const int nwords = (int)(nbytes / 4);
    if (nwords <= 64) {
        PokePayload<64> pl;
        memcpy(pl.w, src_host, (size_t)nbytes);
        bnnp_poke_kernel<64><<<1, 64, 0, (cudaStream_t)stream>>>((uint32_t*)dst, pl, nwords);
    } else {
        PokePayload<960> pl;
        memcpy(pl.w, src_host, (size_t)nbytes);
        bnnp_poke_kernel<960><<<1, 256, 0, (cudaStream_t)stream>>>((uint32_t*)dst, pl, nwords);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "bnnp_poke_kernel launch");
    return 0;
}

int bnnp_probe_stream(float* P, const float* G, float* M, int64_t total, void* stream) {
    if (P == nullptr || G == nullptr || M == nullptr || total <= 0 || total % 4 != 0)
        return fail(BNNP_E_ARG, "bnnp_probe_stream: bad argument");
    const int64_t nquads = total / 4;
    const int64_t blocks = (nquads + THREADS * UNROLL - 1) / (THREADS * UNROLL);
    bnnp_probe_stream_kernel<<<(unsigned)blocks, THREADS, 0, (cudaStream_t)stream>>>(P, G, M, nquads);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "bnnp_probe_stream_kernel launch");
    return 0;
}

int bnnp_rollback(float* P, float* G, float* M, const float* prev_p, const float* prev_g, const float* prev_m,
                  int64_t total, void* stream) {
    if (P == nullptr || G == nullptr || prev_p == nullptr || prev_g == nullptr || total <= 0)
        return fail(BNNP_E_ARG, "bnnp_rollback: null argument");
    if ((prev_m != nullptr) != (M != nullptr) && prev_m != nullptr)
        return fail(BNNP_E_ARG, "bnnp_rollback: prev_m without M");
    if (total % 4 != 0) return fail(BNNP_E_ALIGN, "bnnp_rollback: total must be a multiple of 4 floats");
    if (misaligned(P) || misaligned(G) || misaligned(M) || misaligned(prev_p) || misaligned(prev_g) || misaligned(prev_m))
        return fail(BNNP_E_ALIGN, "bnnp_rollback: flat arrays must be 16-byte aligned");
    const int64_t nquads = total / 4;
    int64_t blocks = (nquads + THREADS - 1) / THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    bnnp_rollback_kernel<<<(int)blocks, THREADS, 0, (cudaStream_t)stream>>>(P, G, M, prev_p, prev_g,
                                                                              M != nullptr ? prev_m : nullptr, nquads);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "bnnp_rollback_kernel launch");
    return 0;
}

}  // extern "C"
#endif  // BNNP_PART == -1 || BNNP_PART == 0
