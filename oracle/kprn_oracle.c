/*
 * kprn_oracle.c -- CPU restatement (fp64) of the KPRN path-scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kprn_b200/ may link, import or call this
 * file.  It is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker and the CPU baseline.
 *
 * PARITY UNPINNED: the reference (eBay/KPRN, Lua/Torch7) cannot run in this image (no
 * lua/luajit/th, no Torch7 libs) and ships no golden vectors or numeric tests for this
 * path.  The arithmetic of the path lives in third-party,
 * un-vendored, un-pinned rocks (torch7, nn, Element-Research rnn, optim); this file
 * restates their published algorithms and anchors on the reference's own call sites,
 * cited per function below (paths relative to /root/reference/release/songPathRnn/).
 * It is cross-checked against an independent PyTorch fp64 autograd model in
 * tests/test_oracle.py.
 *
 * All arithmetic is fp64: Torch7's CPU default tensor is DoubleTensor and
 * model/OneModel.lua never changes it.
 *
 * Flat parameter order (nn.Module:parameters() DFS of model/OneModel.lua:203-276):
 *   E_type[V_t,d_t]  E_entity[V_e,d_e]  E_relation[V_r,d_r]
 *   LSTM : Wx[4H,D] bx[4H] Wh[4H,H]            (D = d_t+d_e+d_r, FastLSTM i2g / o2g)
 *   RNN  : Wx[H,D]  bx[H]  Wh[H,H]  bh[H]      (nn.Linear(D,H), nn.Linear(H,H))
 *   head : Wo[L,H] bo[L]   (2-layer head: W1[H,H] b1[H] W2[L,H] b2[L])
 * ids are [N,T,3] int32, 1-based, column order (type, entity, relation)
 * (data/movie_data_format.py:142-157, data/int2torch.lua:58-63).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t T, d_t, d_e, d_r, H, L;
    int64_t V_t, V_e, V_r;
    int32_t cell;        /* 0 = FastLSTM (OneModel.lua:236), 1 = ReLU Elman RNN + MaskZero (OneModel.lua:240-254) */
    int32_t head_layers; /* 1 = nn.Linear(H,L) (OneModel.lua:275); 2 = Linear(H,H)-ReLU-Linear(H,L) (paper head) */
    int32_t pool;        /* 0 = LogSumExp (module/LogSumExp.lua), 1 = max (nn.Max(2)), 2 = TopK mean (module/TopK.lua) */
    int32_t topk;        /* k for pool==2 */
    double inv_gamma;    /* LSE temperature: lse = log(sum exp(s*inv_gamma))/inv_gamma ; reference = 1.0 */
} okprn_cfg;

typedef struct {
    int64_t E_t, E_e, E_r, Wx, bx, Wh, bh, Wo, bo, W1, b1, total;
    int32_t G; /* gate rows: 4H (lstm) or H (rnn) */
    int32_t D;
} okprn_layout;

static void layout_of(const okprn_cfg* c, okprn_layout* l) {
    int64_t o = 0;
    l->D = c->d_t + c->d_e + c->d_r;
    l->G = (c->cell == 0 ? 4 : 1) * c->H;
    l->E_t = o; o += c->V_t * c->d_t;
    l->E_e = o; o += c->V_e * c->d_e;
    l->E_r = o; o += c->V_r * c->d_r;
    l->Wx = o; o += (int64_t)l->G * l->D;
    l->bx = o; o += l->G;
    l->Wh = o; o += (int64_t)l->G * c->H;
    l->bh = -1;
    if (c->cell == 1) { l->bh = o; o += c->H; }
    l->W1 = l->b1 = -1;
    if (c->head_layers == 2) { l->W1 = o; o += (int64_t)c->H * c->H; l->b1 = o; o += c->H; }
    l->Wo = o; o += (int64_t)c->L * c->H;
    l->bo = o; o += c->L;
    l->total = o;
}

int64_t okprn_param_count(const okprn_cfg* c) { okprn_layout l; layout_of(c, &l); return l.total; }

/* offsets out[11]: E_t,E_e,E_r,Wx,bx,Wh,bh,W1,b1,Wo,bo (-1 when absent) */
void okprn_param_offsets(const okprn_cfg* c, int64_t* out) {
    okprn_layout l; layout_of(c, &l);
    out[0]=l.E_t; out[1]=l.E_e; out[2]=l.E_r; out[3]=l.Wx; out[4]=l.bx; out[5]=l.Wh;
    out[6]=l.bh; out[7]=l.W1; out[8]=l.b1; out[9]=l.Wo; out[10]=l.bo;
}

int okprn_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank: the timed CPU baseline sets its thread count explicitly */
void okprn_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline double sigm(double x) { return 1.0 / (1.0 + exp(-x)); }

/* C[m][n] (+)= sum_k A[m][k] * B[n][k]   (y = x W^T, nn.Linear semantics: W is [out,in]) */
static void gemm_nt(int M, int Nn, int K, const double* A, int lda, const double* B, int ldb,
                    double* C, int ldc, int accumulate) {
    for (int m = 0; m < M; ++m) {
        const double* a = A + (size_t)m * lda;
        for (int n = 0; n < Nn; ++n) {
            const double* b = B + (size_t)n * ldb;
            double acc = 0.0;
#pragma omp simd reduction(+ : acc)
            for (int k = 0; k < K; ++k) acc += a[k] * b[k];
            if (accumulate) C[(size_t)m * ldc + n] += acc; else C[(size_t)m * ldc + n] = acc;
        }
    }
}
/* C[m][n] += sum_k A[m][k] * B[k][n] */
static void gemm_nn_acc(int M, int Nn, int K, const double* A, int lda, const double* B, int ldb,
                        double* C, int ldc) {
    for (int m = 0; m < M; ++m) {
        double* c = C + (size_t)m * ldc;
        for (int k = 0; k < K; ++k) {
            const double a = A[(size_t)m * lda + k];
            const double* b = B + (size_t)k * ldb;
#pragma omp simd
            for (int n = 0; n < Nn; ++n) c[n] += a * b[n];
        }
    }
}
/* C[i][j] += sum_m A[m][i] * B[m][j]   (weight gradient: gW += dG^T x) */
static void gemm_tn_acc(int M, int I, int J, const double* A, int lda, const double* B, int ldb,
                        double* C, int ldc) {
    for (int m = 0; m < M; ++m) {
        for (int i = 0; i < I; ++i) {
            const double a = A[(size_t)m * lda + i];
            if (a == 0.0) continue;
            const double* b = B + (size_t)m * ldb;
            double* c = C + (size_t)i * ldc;
#pragma omp simd
            for (int j = 0; j < J; ++j) c[j] += a * b[j];
        }
    }
}

#define PB 16 /* paths per block */

/* model/net/FeatureEmbedding.lua:112-121 -- concat order type . entity . relation
 * (three nn.LookupTable row selects + JoinTable(3)). Returns nonzero on a bad id. */
static int gather_x(const okprn_cfg* c, const okprn_layout* l, const double* P, const int32_t* ids,
                    int64_t n0, int nb, int t, double* x /*[nb,D]*/) {
    for (int p = 0; p < nb; ++p) {
        const int32_t* id = ids + ((n0 + p) * c->T + t) * 3;
        int64_t it = id[0] - 1, ie = id[1] - 1, ir = id[2] - 1;
        if (it < 0 || it >= c->V_t || ie < 0 || ie >= c->V_e || ir < 0 || ir >= c->V_r) return 1;
        double* xr = x + (size_t)p * l->D;
        memcpy(xr, P + l->E_t + it * c->d_t, sizeof(double) * c->d_t);
        memcpy(xr + c->d_t, P + l->E_e + ie * c->d_e, sizeof(double) * c->d_e);
        memcpy(xr + c->d_t + c->d_e, P + l->E_r + ir * c->d_r, sizeof(double) * c->d_r);
    }
    return 0;
}

/* MaskZero(rm,1) (OneModel.lua:251): a step whose input row is all-zero outputs zero.
 * With the pad rows of all three tables zeroed (MyOptimizer.lua:74-93) this is the pad step. */
static int row_is_zero(const double* x, int D) {
    for (int k = 0; k < D; ++k) if (x[k] != 0.0) return 0;
    return 1;
}

/*
 * Per-path forward: embeddings -> recurrent cell over all T steps -> head.
 *   LSTM: Element-Research rnn FastLSTM: gates = i2g(x) + o2g(h), block order
 *         [input sigma, hidden(candidate) tanh, forget sigma, output sigma];
 *         c = i*g + f*c_prev ; h = o*tanh(c); h0=c0=0; no masking (OneModel.lua:236,268-273).
 *   head: SelectTable(-1) + nn.Linear(H,L) (OneModel.lua:274-275).
 * cache layout (may be NULL): act[N,T,G] post-activation gates, cst[N,T,H], hst[N,T,H],
 * and for head_layers==2 z1[N,H] (post-ReLU hidden of the head).
 */
int okprn_forward_paths(const okprn_cfg* c, const double* P, const int32_t* ids, int64_t N,
                        double* path_logits /*[N,L]*/, double* act, double* cst, double* hst, double* z1c) {
    okprn_layout l; layout_of(c, &l);
    const int H = c->H, D = l.D, G = l.G, T = c->T, L = c->L;
    int err = 0;
    int64_t nblk = (N + PB - 1) / PB;
#pragma omp parallel
    {
        double* x = (double*)malloc(sizeof(double) * PB * D);
        double* g = (double*)malloc(sizeof(double) * PB * G);
        double* h = (double*)malloc(sizeof(double) * PB * H);
        double* cc = (double*)malloc(sizeof(double) * PB * H);
        double* z1 = (double*)malloc(sizeof(double) * PB * H);
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < nblk; ++b) {
            int64_t n0 = b * PB;
            int nb = (int)((N - n0) < PB ? (N - n0) : PB);
            memset(h, 0, sizeof(double) * PB * H);
            memset(cc, 0, sizeof(double) * PB * H);
            for (int t = 0; t < T; ++t) {
                if (gather_x(c, &l, P, ids, n0, nb, t, x)) { err = 1; continue; }
                for (int p = 0; p < nb; ++p) memcpy(g + (size_t)p * G, P + l.bx, sizeof(double) * G);
                gemm_nt(nb, G, D, x, D, P + l.Wx, D, g, G, 1);
                gemm_nt(nb, G, H, h, H, P + l.Wh, H, g, G, 1);
                for (int p = 0; p < nb; ++p) {
                    double* gp = g + (size_t)p * G;
                    double* hp = h + (size_t)p * H;
                    double* cp = cc + (size_t)p * H;
                    if (c->cell == 0) {
                        for (int u = 0; u < H; ++u) {
                            double i = sigm(gp[u]), gg = tanh(gp[H + u]), f = sigm(gp[2 * H + u]), o = sigm(gp[3 * H + u]);
                            double cn = i * gg + f * cp[u];
                            cp[u] = cn; hp[u] = o * tanh(cn);
                            gp[u] = i; gp[H + u] = gg; gp[2 * H + u] = f; gp[3 * H + u] = o;
                        }
                    } else {
                        int masked = row_is_zero(x + (size_t)p * D, D);
                        for (int u = 0; u < H; ++u) {
                            double v = gp[u] + P[l.bh + u];
                            v = v > 0.0 ? v : 0.0;
                            if (masked) v = 0.0;
                            hp[u] = v; gp[u] = masked ? -1.0 : v; /* act<0 marks a masked step */
                        }
                    }
                    if (act) memcpy(act + ((n0 + p) * T + t) * (size_t)G, gp, sizeof(double) * G);
                    if (cst && c->cell == 0) memcpy(cst + ((n0 + p) * T + t) * (size_t)H, cp, sizeof(double) * H);
                    if (hst) memcpy(hst + ((n0 + p) * T + t) * (size_t)H, hp, sizeof(double) * H);
                }
            }
            const double* hin = h;
            if (c->head_layers == 2) {
                for (int p = 0; p < nb; ++p) memcpy(z1 + (size_t)p * H, P + l.b1, sizeof(double) * H);
                gemm_nt(nb, H, H, h, H, P + l.W1, H, z1, H, 1);
                for (int k = 0; k < nb * H; ++k) z1[k] = z1[k] > 0.0 ? z1[k] : 0.0;
                if (z1c) memcpy(z1c + n0 * (size_t)H, z1, sizeof(double) * nb * H);
                hin = z1;
            }
            for (int p = 0; p < nb; ++p) memcpy(path_logits + (n0 + p) * (size_t)L, P + l.bo, sizeof(double) * L);
            gemm_nt(nb, L, H, hin, H, P + l.Wo, H, path_logits + n0 * (size_t)L, L, 1);
        }
        free(x); free(g); free(h); free(cc); free(z1);
    }
    return err;
}

static int cmp_desc(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x < y) - (x > y);
}

/*
 * Pool the paths of each (user,item) segment, per column.
 *   pool 0: module/LogSumExp.lua:13-30 (m = max; lse = m + log sum exp(s-m)); MapReduce.lua:32-47 views.
 *   pool 1: nn.Max(2) (OneModel.lua:291-292).   pool 2: module/TopK.lua + nn.Mean(2) (OneModel.lua:285-286).
 * seg is CSR offsets [S+1] over consecutive paths; a dense reference bucket [B,P] is seg[k]=k*P.
 * scores[k] = sigmoid(pooled[k, class_id-1])  (nn.Sigmoid + nn.Select(2,classId),
 * OneModel.lua:294, MyOptimizer.lua:126, eval/test_from_checkpoint.lua:81-82).
 */
int okprn_pool(const okprn_cfg* c, const double* path_logits, const int32_t* seg, int64_t S,
               int class_id, double* seg_logits /*[S,L]*/, double* scores /*[S] or NULL*/) {
    const int L = c->L;
    const double ig = c->inv_gamma == 0.0 ? 1.0 : c->inv_gamma;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < S; ++k) {
        int64_t a = seg[k], b = seg[k + 1];
        for (int j = 0; j < L; ++j) {
            double out;
            if (b <= a) { out = 0.0; }
            else if (c->pool == 0) {
                double m = -INFINITY;
                for (int64_t p = a; p < b; ++p) { double v = path_logits[p * L + j] * ig; if (v > m) m = v; }
                double z = 0.0;
                for (int64_t p = a; p < b; ++p) z += exp(path_logits[p * L + j] * ig - m);
                out = (m + log(z)) / ig;
            } else if (c->pool == 1) {
                double m = -INFINITY;
                for (int64_t p = a; p < b; ++p) if (path_logits[p * L + j] > m) m = path_logits[p * L + j];
                out = m;
            } else {
                int64_t n = b - a; int64_t kk = c->topk < n ? c->topk : n;
                double* tmp = (double*)malloc(sizeof(double) * n);
                for (int64_t p = 0; p < n; ++p) tmp[p] = path_logits[(a + p) * L + j];
                qsort(tmp, n, sizeof(double), cmp_desc);
                double s = 0.0; for (int64_t p = 0; p < kk; ++p) s += tmp[p];
                out = s / (double)kk; free(tmp);
            }
            seg_logits[k * L + j] = out;
        }
        if (scores) scores[k] = sigm(seg_logits[k * L + (class_id - 1)]);
    }
    return 0;
}

/* nn.BCECriterion (torch/nn, eps=1e-12, sizeAverage): OneModel.lua:300, MyOptimizer.lua:193-194.
 * loss = -(1/S) sum [ y log(p+eps) + (1-y) log(1-p+eps) ];  dL/dp = -(y-p)/((1-p+eps)(p+eps))/S */
double okprn_bce(const double* p, const double* y, int64_t S, double* dp) {
    const double eps = 1e-12; double loss = 0.0;
    for (int64_t k = 0; k < S; ++k) {
        loss -= y[k] * log(p[k] + eps) + (1.0 - y[k]) * log(1.0 - p[k] + eps);
        if (dp) dp[k] = -(y[k] - p[k]) / ((1.0 - p[k] + eps) * (p[k] + eps)) / (double)S;
    }
    return loss / (double)S;
}

/* model/criterion/BPRLoss.lua:18-32: theta=pos-neg; out=sigmoid(theta)+1e-4; err=-mean(log out);
 * gradInput = { -(1-out), +(1-out) }  (NOT divided by n; eps folded in -- reproduced exactly). */
double okprn_bpr(const double* pos, const double* neg, int64_t n, double* dpos, double* dneg) {
    double loss = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        double out = sigm(pos[k] - neg[k]) + 1e-4;
        loss -= log(out);
        double step = 1.0 - out;
        if (dpos) dpos[k] = -step;
        if (dneg) dneg[k] = step;
    }
    return loss / (double)n;
}

/* gradient of scores[k]=sigmoid(seg_logits[k,c]) back to seg_logits: nn.Select + nn.Sigmoid backward */
void okprn_score_backward(const okprn_cfg* c, const double* scores, const double* dscores, int64_t S,
                          int class_id, double* dseg_logits /*[S,L] zero-filled then set*/) {
    memset(dseg_logits, 0, sizeof(double) * S * c->L);
    for (int64_t k = 0; k < S; ++k)
        dseg_logits[k * c->L + (class_id - 1)] = dscores[k] * scores[k] * (1.0 - scores[k]);
}

/*
 * Backward of pool + head + BPTT + embedding scatter-add; accumulates into grads (flat, same layout).
 *   LogSumExp.lua:32-37 (softmax * gradOutput), MapReduce.lua:74-85, nn.Linear / FastLSTM BPTT /
 *   LookupTable:accGradParameters (scatter-add of gradOutput rows). SURVEY.md App. A.7.
 * Pad rows receive gradient too (they are re-zeroed in the weights by zeroPadTokens, not in the grads).
 */
int okprn_backward(const okprn_cfg* c, const double* P, const int32_t* ids, const int32_t* seg,
                   int64_t N, int64_t S, const double* path_logits, const double* seg_logits,
                   const double* dseg_logits, const double* act, const double* cst, const double* hst,
                   const double* z1c, double scale, double* grads) {
    okprn_layout l; layout_of(c, &l);
    const int H = c->H, D = l.D, G = l.G, T = c->T, L = c->L;
    const double ig = c->inv_gamma == 0.0 ? 1.0 : c->inv_gamma;
    double* ds = (double*)calloc((size_t)N * L, sizeof(double));
    /* pool backward */
    for (int64_t k = 0; k < S; ++k) {
        int64_t a = seg[k], b = seg[k + 1];
        if (b <= a) continue;
        for (int j = 0; j < L; ++j) {
            double go = dseg_logits[k * L + j] * scale;
            if (go == 0.0) continue;
            if (c->pool == 0) {
                for (int64_t p = a; p < b; ++p)
                    ds[p * L + j] = exp((path_logits[p * L + j] - seg_logits[k * L + j]) * ig) * go;
            } else if (c->pool == 1) {
                int64_t best = a;
                for (int64_t p = a; p < b; ++p) if (path_logits[p * L + j] > path_logits[best * L + j]) best = p;
                ds[best * L + j] = go;
            } else {
                int64_t n = b - a; int64_t kk = c->topk < n ? c->topk : n;
                /* top-k (stable: earlier index wins ties), mean backward */
                char* used = (char*)calloc(n, 1);
                for (int64_t r = 0; r < kk; ++r) {
                    int64_t best = -1;
                    for (int64_t p = 0; p < n; ++p) if (!used[p] && (best < 0 || path_logits[(a + p) * L + j] > path_logits[(a + best) * L + j])) best = p;
                    used[best] = 1; ds[(a + best) * L + j] = go / (double)kk;
                }
                free(used);
            }
        }
    }
    int err = 0;
    int nthreads = okprn_num_threads();
    size_t dense_n = (size_t)(l.total - l.Wx); /* Wx..bo are contiguous at the tail */
    double* tl = (double*)calloc(dense_n * nthreads, sizeof(double));
    int64_t nblk = (N + PB - 1) / PB;
#pragma omp parallel
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double* gd = tl + dense_n * tid - l.Wx; /* gd[l.Wx + i] addresses the thread-local dense grads */
        double* x = (double*)malloc(sizeof(double) * PB * D);
        double* dG = (double*)malloc(sizeof(double) * PB * G);
        double* dh = (double*)malloc(sizeof(double) * PB * H);
        double* dc = (double*)malloc(sizeof(double) * PB * H);
        double* dx = (double*)malloc(sizeof(double) * PB * D);
        double* dz = (double*)malloc(sizeof(double) * PB * H);
        double* hprev = (double*)malloc(sizeof(double) * PB * H);
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < nblk; ++b) {
            int64_t n0 = b * PB;
            int nb = (int)((N - n0) < PB ? (N - n0) : PB);
            const double* dsb = ds + n0 * (size_t)L;
            /* head backward */
            for (int p = 0; p < nb; ++p) {
                const double* hT = hst + ((n0 + p) * T + (T - 1)) * (size_t)H;
                const double* hin = (c->head_layers == 2) ? z1c + (n0 + p) * (size_t)H : hT;
                for (int j = 0; j < L; ++j) {
                    double g = dsb[p * L + j];
                    if (g == 0.0) continue;
                    gd[l.bo + j] += g;
                    for (int u = 0; u < H; ++u) gd[l.Wo + (size_t)j * H + u] += g * hin[u];
                }
            }
            memset(dh, 0, sizeof(double) * PB * H);
            if (c->head_layers == 2) {
                memset(dz, 0, sizeof(double) * PB * H);
                gemm_nn_acc(nb, H, L, dsb, L, P + l.Wo, H, dz, H);
                for (int p = 0; p < nb; ++p) {
                    const double* z = z1c + (n0 + p) * (size_t)H;
                    const double* hT = hst + ((n0 + p) * T + (T - 1)) * (size_t)H;
                    for (int u = 0; u < H; ++u) {
                        double g = z[u] > 0.0 ? dz[p * H + u] : 0.0;
                        dz[p * H + u] = g;
                        gd[l.b1 + u] += g;
                        for (int k = 0; k < H; ++k) gd[l.W1 + (size_t)u * H + k] += g * hT[k];
                    }
                }
                gemm_nn_acc(nb, H, H, dz, H, P + l.W1, H, dh, H);
            } else {
                gemm_nn_acc(nb, H, L, dsb, L, P + l.Wo, H, dh, H);
            }
            memset(dc, 0, sizeof(double) * PB * H);
            for (int t = T - 1; t >= 0; --t) {
                if (gather_x(c, &l, P, ids, n0, nb, t, x)) { err = 1; break; }
                for (int p = 0; p < nb; ++p) {
                    const double* a = act + ((n0 + p) * T + t) * (size_t)G;
                    double* dGp = dG + (size_t)p * G;
                    double* dhp = dh + (size_t)p * H;
                    double* hp = hprev + (size_t)p * H;
                    if (t > 0) memcpy(hp, hst + ((n0 + p) * T + t - 1) * (size_t)H, sizeof(double) * H);
                    else memset(hp, 0, sizeof(double) * H);
                    if (c->cell == 0) {
                        const double* ct = cst + ((n0 + p) * T + t) * (size_t)H;
                        const double* cprev = t > 0 ? cst + ((n0 + p) * T + t - 1) * (size_t)H : NULL;
                        double* dcp = dc + (size_t)p * H;
                        for (int u = 0; u < H; ++u) {
                            double i = a[u], gg = a[H + u], f = a[2 * H + u], o = a[3 * H + u];
                            double tc = tanh(ct[u]);
                            double d_o = dhp[u] * tc;
                            double dcu = dcp[u] + dhp[u] * o * (1.0 - tc * tc);
                            double di = dcu * gg, dg = dcu * i, df = dcu * (cprev ? cprev[u] : 0.0);
                            dcp[u] = dcu * f;
                            dGp[u] = di * i * (1.0 - i);
                            dGp[H + u] = dg * (1.0 - gg * gg);
                            dGp[2 * H + u] = df * f * (1.0 - f);
                            dGp[3 * H + u] = d_o * o * (1.0 - o);
                        }
                    } else {
                        for (int u = 0; u < H; ++u) dGp[u] = (a[u] > 0.0) ? dhp[u] : 0.0; /* ReLU'; masked step (act<0) -> 0 */
                        for (int u = 0; u < H; ++u) gd[l.bh + u] += dGp[u];
                    }
                    for (int u = 0; u < G; ++u) gd[l.bx + u] += dGp[u];
                }
                gemm_tn_acc(nb, G, D, dG, G, x, D, gd + l.Wx, D);
                gemm_tn_acc(nb, G, H, dG, G, hprev, H, gd + l.Wh, H);
                memset(dx, 0, sizeof(double) * PB * D);
                gemm_nn_acc(nb, D, G, dG, G, P + l.Wx, D, dx, D);
                memset(dh, 0, sizeof(double) * PB * H);
                gemm_nn_acc(nb, H, G, dG, G, P + l.Wh, H, dh, H);
                /* LookupTable:accGradParameters -- scatter-add */
                for (int p = 0; p < nb; ++p) {
                    const int32_t* id = ids + ((n0 + p) * T + t) * 3;
                    const double* dxp = dx + (size_t)p * D;
                    double* gt = grads + l.E_t + (int64_t)(id[0] - 1) * c->d_t;
                    double* ge = grads + l.E_e + (int64_t)(id[1] - 1) * c->d_e;
                    double* gr = grads + l.E_r + (int64_t)(id[2] - 1) * c->d_r;
                    for (int k = 0; k < c->d_t; ++k) {
#pragma omp atomic
                        gt[k] += dxp[k];
                    }
                    for (int k = 0; k < c->d_e; ++k) {
#pragma omp atomic
                        ge[k] += dxp[c->d_t + k];
                    }
                    for (int k = 0; k < c->d_r; ++k) {
#pragma omp atomic
                        gr[k] += dxp[c->d_t + c->d_e + k];
                    }
                }
            }
        }
        free(x); free(dG); free(dh); free(dc); free(dx); free(dz); free(hprev);
    }
    for (int t = 0; t < nthreads; ++t)
        for (size_t i = 0; i < dense_n; ++i) grads[l.Wx + i] += tl[dense_n * t + i];
    free(tl); free(ds);
    return err;
}

/* MyOptimizer.lua:196-214: if regularize: clip global norm to clipnorm (when useGradClip), then g += l2*theta */
void okprn_clip_l2(double* g, const double* theta, int64_t n, int use_clip, double clipnorm, double l2) {
    if (use_clip) {
        double s = 0.0; for (int64_t i = 0; i < n; ++i) s += g[i] * g[i];
        double nrm = sqrt(s);
        if (nrm > clipnorm) { double f = clipnorm / nrm; for (int64_t i = 0; i < n; ++i) g[i] *= f; }
    }
    if (l2 != 0.0) for (int64_t i = 0; i < n; ++i) g[i] += l2 * theta[i];
}

/* optim.adam (torch/optim): OneModel.lua:340-355 (lr 1e-3, beta1 .9, beta2 .999, eps 1e-8).
 * m=b1*m+(1-b1)g; v=b2*v+(1-b2)g^2; denom=sqrt(v)+eps; step=lr*sqrt(1-b2^t)/(1-b1^t); x-=step*m/denom.
 * t is the 1-based step count AFTER increment. Dense over every parameter (incl. all entity rows). */
void okprn_adam(double* x, const double* g, double* m, double* v, int64_t n, int64_t t,
                double lr, double b1, double b2, double eps) {
    double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
    double step = lr * sqrt(bc2) / bc1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        m[i] = b1 * m[i] + (1.0 - b1) * g[i];
        v[i] = b2 * v[i] + (1.0 - b2) * g[i] * g[i];
        x[i] -= step * m[i] / (sqrt(v[i]) + eps);
    }
}

/* MyOptimizer.lua:74-93 zeroPadTokens: params[1][idx]:zero(); table 0 type, 1 entity, 2 relation; row 1-based */
int okprn_zero_row(const okprn_cfg* c, double* P, int table, int64_t row_1based) {
    okprn_layout l; layout_of(c, &l);
    int64_t V = table == 0 ? c->V_t : table == 1 ? c->V_e : c->V_r;
    int d = table == 0 ? c->d_t : table == 1 ? c->d_e : c->d_r;
    int64_t off = table == 0 ? l.E_t : table == 1 ? l.E_e : l.E_r;
    if (row_1based < 1 || row_1based > V) return 1;
    memset(P + off + (row_1based - 1) * d, 0, sizeof(double) * d);
    return 0;
}
