/* A consumer of the C ABI (include/bnnp.h) that is neither Python nor torch: plain C + the
 * CUDA runtime.  It lays out a chain of three ragged tensors with bnnp_plan_layout, runs one
 * SGLD transition with momentum 0 and temperature 0 -- the SGD known answer of the reference's
 * test_sgd_equivalence (testing/test_sgld.py:61-80): p' = p - lr * g -- then a second
 * transition that carries the first one's deferred epilogue, finalises and reads the
 * per-tensor dot(g, g) back.  Built and run by tests/test_cuda_c_abi.py.
 *
 *   gcc tests/c/abi_consumer.c -I include -I /usr/local/cuda/include -L bnn_priors_b200/_lib -lbnnp \
 *       -L /usr/local/cuda/lib64 -lcudart -lm -o abi_consumer
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "bnnp.h"
#include "bnnp_eval.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_BNNP(x) do { int rc_ = (x); if (rc_ != 0) { \
    fprintf(stderr, "%s:%d rc=%d %s\n", __FILE__, __LINE__, rc_, bnnp_last_error()); return 3; } } while (0)

int main(void) {
    enum { NSEG = 3 };
    const int64_t numel[NSEG] = {10, 5000, 4097};
    int64_t off[NSEG], total = 0;
    int32_t first[NSEG], nch[NSEG], nchunks = 0;
    if (bnnp_abi_version() != BNNP_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    CHECK_BNNP(bnnp_plan_layout(numel, NSEG, off, first, nch, &total, &nchunks, NULL));
    BnnpChunk* chunks = (BnnpChunk*)calloc((size_t)nchunks, sizeof(BnnpChunk));
    CHECK_BNNP(bnnp_plan_layout(numel, NSEG, off, first, nch, &total, &nchunks, chunks));
    if (nchunks != 1 + 2 + 2 || total % BNNP_SEG_ALIGN != 0) { fprintf(stderr, "unexpected plan\n"); return 1; }

    /* host copies: p and g in the flat layout, padding zero */
    float* hp = (float*)calloc((size_t)total, sizeof(float));
    float* hg = (float*)calloc((size_t)total, sizeof(float));
    double gg_want[NSEG] = {0, 0, 0};
    unsigned s = 12345u;
    for (int t = 0; t < NSEG; ++t)
        for (int64_t i = 0; i < numel[t]; ++i) {
            s = s * 1664525u + 1013904223u;
            hp[off[t] + i] = (float)((s >> 8) & 0xffff) / 65536.0f - 0.5f;
            s = s * 1664525u + 1013904223u;
            hg[off[t] + i] = ((float)((s >> 8) & 0xffff) / 65536.0f - 0.5f) * 0.1f;
            gg_want[t] += (double)hg[off[t] + i] * (double)hg[off[t] + i];
        }

    BnnpSegment segs[NSEG];
    memset(segs, 0, sizeof(segs));
    for (int t = 0; t < NSEG; ++t) {
        segs[t].off = off[t]; segs[t].numel = numel[t]; segs[t].precond = 1.0;
        segs[t].prior_kind = BNNP_PRIOR_NONE; segs[t].prior_scale = 1.0f; segs[t].prior_df = 3.0f;
        segs[t].first_chunk = first[t]; segs[t].num_chunks = nch[t]; segs[t].link = -1;
    }
    double hstate[NSEG][BNNP_STATE_STRIDE];
    memset(hstate, 0, sizeof(hstate));
    for (int t = 0; t < NSEG; ++t) hstate[t][BNNP_S_SQ_MEAN] = 1.0;

    float *P, *G;
    BnnpSegment* dsegs; BnnpChunk* dchunks; double *dstate, *partials; uint64_t* stamps;
    CHECK_CUDA(cudaMalloc((void**)&P, (size_t)total * 4));
    CHECK_CUDA(cudaMalloc((void**)&G, (size_t)total * 4));
    CHECK_CUDA(cudaMalloc((void**)&dsegs, sizeof(segs)));
    CHECK_CUDA(cudaMalloc((void**)&dchunks, (size_t)nchunks * sizeof(BnnpChunk)));
    CHECK_CUDA(cudaMalloc((void**)&dstate, sizeof(hstate)));
    CHECK_CUDA(cudaMalloc((void**)&partials, (size_t)2 * nchunks * BNNP_NRED * sizeof(double)));
    CHECK_CUDA(cudaMalloc((void**)&stamps, (size_t)2 * nchunks * sizeof(uint64_t)));
    CHECK_CUDA(cudaMemcpy(P, hp, (size_t)total * 4, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(G, hg, (size_t)total * 4, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dsegs, segs, sizeof(segs), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dchunks, chunks, (size_t)nchunks * sizeof(BnnpChunk), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dstate, hstate, sizeof(hstate), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemset(partials, 0, (size_t)2 * nchunks * BNNP_NRED * sizeof(double)));
    CHECK_CUDA(cudaMemset(stamps, 0, (size_t)2 * nchunks * sizeof(uint64_t)));

    /* SGLD, momentum 0, temperature 0 (sgld.py:114-154): m = -hn g, p += h m, hn h = lr */
    const double lr = 0.25, N = 4.0;
    BnnpLaunch a;
    memset(&a, 0, sizeof(a));
    a.P = P; a.G = G; a.segs = dsegs; a.chunks = dchunks; a.seg_state = dstate; a.partials = partials; a.stamps = stamps;
    a.nseg = NSEG; a.nchunks = nchunks; a.nchunks_total = nchunks;
    a.op = BNNP_OP_SGLD; a.phase = BNNP_PHASE_MID; a.noise = BNNP_NOISE_NONE;
    a.flags = BNNP_F_READ_P | BNNP_F_READ_G | BNNP_F_WRITE_P | BNNP_F_UPDATE_SQ | BNNP_F_MM_PRE_NOISE;
    a.cm = 0.0; a.cg = -sqrt(lr * N); a.cn = 0.0; a.cp = sqrt(lr / N); a.rms_alpha = 0.99;
    a.parity = 0; a.call = 0; a.pending.valid = 0;
    CHECK_BNNP(bnnp_launch(&a, NULL));
    /* the second launch carries the first one's epilogue (include/bnnp.h: BnnpEpilogue) */
    BnnpEpilogue pend;
    memset(&pend, 0, sizeof(pend));
    pend.valid = 1; pend.op = a.op; pend.phase = a.phase; pend.flags = a.flags; pend.parity = 0; pend.call = 0;
    pend.rms_alpha = a.rms_alpha;
    a.pending = pend; a.parity = 1; a.call = 1; a.flags |= BNNP_F_REVERSE;
    CHECK_BNNP(bnnp_launch(&a, NULL));
    pend.parity = 1; pend.call = 1; pend.flags = a.flags;
    a.pending = pend;
    CHECK_BNNP(bnnp_finalize(&a, NULL));
    CHECK_CUDA(cudaDeviceSynchronize());

    float* got = (float*)malloc((size_t)total * 4);
    CHECK_CUDA(cudaMemcpy(got, P, (size_t)total * 4, cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(hstate, dstate, sizeof(hstate), cudaMemcpyDeviceToHost));
    double worst = 0.0;
    for (int t = 0; t < NSEG; ++t)
        for (int64_t i = 0; i < numel[t]; ++i) {
            const double want = (double)hp[off[t] + i] - 2.0 * lr * (double)hg[off[t] + i];   /* two steps */
            const double d = fabs((double)got[off[t] + i] - want);
            if (d > worst) worst = d;
        }
    for (int64_t i = 0; i < total; ++i) {           /* padding untouched */
        int inside = 0;
        for (int t = 0; t < NSEG; ++t) inside |= (i >= off[t] && i < off[t] + numel[t]);
        if (!inside && got[i] != 0.0f) { fprintf(stderr, "padding written at %lld\n", (long long)i); return 1; }
    }
    if (worst > 2e-7) { fprintf(stderr, "SGD known answer off by %g\n", worst); return 1; }
    for (int t = 0; t < NSEG; ++t) {
        const double gg = hstate[t][BNNP_S_SUM_GG], sq = hstate[t][BNNP_S_SQ_MEAN];
        const double sq_want = 0.99 * (0.99 * 1.0 + 0.01 * gg_want[t] / (double)numel[t]) + 0.01 * gg_want[t] / (double)numel[t];
        if (fabs(gg - gg_want[t]) > 1e-6 * gg_want[t] || hstate[t][BNNP_S_LAUNCHES] != 2.0 ||
            fabs(sq - sq_want) > 1e-9 || hstate[t][BNNP_S_NONFINITE] != 0.0) {
            fprintf(stderr, "segment %d: gg %g want %g, sq %g want %g, launches %g\n", t, gg, gg_want[t], sq, sq_want,
                    hstate[t][BNNP_S_LAUNCHES]);
            return 1;
        }
    }
    /* argument validation reports through the return code and bnnp_last_error */
    a.pending.valid = 0; a.op = 99;
    if (bnnp_launch(&a, NULL) != BNNP_E_ARG || strstr(bnnp_last_error(), "bad op") == NULL) return 1;
    /* include/bnnp_eval.h is plain C too; a zeroed state is rejected before any CUDA call */
    BnnpEvalState ev;
    memset(&ev, 0, sizeof(ev));
    if (bnnp_eval_finish(&ev, NULL, NULL, 1, NULL, NULL, NULL) != BNNP_E_ARG ||
        strstr(bnnp_eval_last_error(), "bad state") == NULL) return 1;
    printf("abi_consumer ok: %d chunks, max |p - (p0 - 2 lr g)| = %.3g\n", nchunks, worst);
    return 0;
}
