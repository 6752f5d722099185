/*
 * oracle/dag_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, never linked into the product).
 *
 * Plain-C restatement of the semantics of the reference CUDA operators of
 * ictnlp/DASpeech, DASpeech/custom_ops (citations are reference file:line):
 *   alpha recurrence        dag_loss.cu:71-131
 *   beta recurrence         dag_loss.cu:206-265
 *   emission gradient       dag_loss.cu:395-399
 *   transition gradient     dag_loss.cu:461-484
 *   Viterbi + back-pointers dag_best_alignment.cu:72-122
 *   backtrace               dag_best_alignment.cu:178-184
 *   logsoftmax + gather     logsoftmax_gather.cu:268-308
 *   gather backward         dag_loss.py:293-295
 *
 * This header is included twice by dag_oracle.c, once with REAL=float (suffix _f32:
 * mimics the reference's fp32 arithmetic; candidate sums for the Viterbi are single
 * IEEE fp32 adds so arg-max indices are reproducible bit for bit) and once with
 * REAL=double (suffix _f64: the "true value" used for tolerance checks).
 *
 * Layouts (all row-major, contiguous):
 *   match [B][M][L]   links [B][L][T]  (links[b][i][k] = log P(i -> i+k+1))
 *   alpha/beta/grad_match [B][M][L]    grad_links [B][L][T]
 *   olen/tlen [B] int64
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

static inline int FN(is_inf)(REAL x) { return isinf(x); }

/* log-sum-exp rule shared by alpha and beta (dag_loss.cu:94-127 / 230-261):
 * mx = max(S); if mx is +-inf the result is mx itself and NO emission is added. */

void FN(oracle_alpha)(const REAL *match, const REAL *links, const int64_t *olen, const int64_t *tlen,
                      REAL *alpha, int B, int M, int L, int T)
{
    const REAL NI = (REAL)(-INFINITY);
    for (int64_t x = 0; x < (int64_t)B * M * L; x++) alpha[x] = NI;
    for (int b = 0; b < B; b++) {
        const int O = (int)olen[b], Tn = (int)tlen[b];
        const REAL *m = match + (int64_t)b * M * L;
        const REAL *E = links + (int64_t)b * L * T;
        REAL *a = alpha + (int64_t)b * M * L;
        a[0] = m[0];
        for (int t = 1; t < Tn; t++) {
            const REAL *ap = a + (int64_t)(t - 1) * L;
            REAL *an = a + (int64_t)t * L;
#pragma omp parallel for schedule(static) if (O - t > 128)
            for (int j = t; j < O; j++) {
                int maxdelta = j < T ? j : T;
                REAL mx = NI;
                for (int d = 1; d <= maxdelta; d++) {
                    REAL v = ap[j - d] + E[(int64_t)(j - d) * T + (d - 1)];
                    if (v > mx) mx = v;
                }
                if (FN(is_inf)(mx)) { an[j] = mx; continue; }
                REAL s = 0;
                for (int d = 1; d <= maxdelta; d++) {
                    REAL v = ap[j - d] + E[(int64_t)(j - d) * T + (d - 1)];
                    s += EXPF(v - mx);
                }
                an[j] = LOGF(s) + mx + m[(int64_t)t * L + j];
            }
        }
    }
}

void FN(oracle_beta)(const REAL *match, const REAL *links, const int64_t *olen, const int64_t *tlen,
                     REAL *beta, int B, int M, int L, int T)
{
    const REAL NI = (REAL)(-INFINITY);
    for (int64_t x = 0; x < (int64_t)B * M * L; x++) beta[x] = NI;
    for (int b = 0; b < B; b++) {
        const int O = (int)olen[b], Tn = (int)tlen[b];
        const REAL *m = match + (int64_t)b * M * L;
        const REAL *E = links + (int64_t)b * L * T;
        REAL *be = beta + (int64_t)b * M * L;
        be[(int64_t)(Tn - 1) * L + (O - 1)] = m[(int64_t)(Tn - 1) * L + (O - 1)];
        for (int t = Tn - 2; t >= 0; t--) {
            const REAL *bn = be + (int64_t)(t + 1) * L;
            REAL *bc = be + (int64_t)t * L;
#pragma omp parallel for schedule(static) if (O - t > 128)
            for (int j = t; j < O; j++) {
                int maxdelta = (O - 1 - j) < T ? (O - 1 - j) : T;
                const REAL *Ej = E + (int64_t)j * T;
                REAL mx = NI;
                for (int d = 1; d <= maxdelta; d++) {
                    REAL v = bn[j + d] + Ej[d - 1];
                    if (v > mx) mx = v;
                }
                if (FN(is_inf)(mx)) { bc[j] = mx; continue; }
                REAL s = 0;
                for (int d = 1; d <= maxdelta; d++) s += EXPF(bn[j + d] + Ej[d - 1] - mx);
                bc[j] = LOGF(s) + mx + m[(int64_t)t * L + j];
            }
        }
    }
}

/* dag_loss.py:107-110 : Z = beta[b,0,0] when a gradient is required, else alpha[b,Tn-1,O-1] */
void FN(oracle_loss)(const REAL *alpha, const REAL *beta, const int64_t *olen, const int64_t *tlen,
                     REAL *loss, int B, int M, int L, int require_gradient)
{
    for (int b = 0; b < B; b++) {
        if (require_gradient) loss[b] = beta[(int64_t)b * M * L];
        else loss[b] = alpha[(int64_t)b * M * L + (int64_t)(tlen[b] - 1) * L + (olen[b] - 1)];
    }
}

void FN(oracle_grad_match)(const REAL *go, const REAL *alpha, const REAL *beta, const REAL *match,
                           REAL *gm, int B, int M, int L)
{
    for (int b = 0; b < B; b++) {
        const int64_t base = (int64_t)b * M * L;
        const REAL Z = beta[base];
#pragma omp parallel for schedule(static)
        for (int64_t x = 0; x < (int64_t)M * L; x++) {
            REAL mm = match[base + x];
            if (FN(is_inf)(mm) || FN(is_inf)(Z)) gm[base + x] = 0;
            else gm[base + x] = EXPF(alpha[base + x] + beta[base + x] - mm - Z) * go[b];
        }
    }
}

void FN(oracle_grad_links)(const REAL *go, const REAL *alpha, const REAL *beta, const REAL *links,
                           const int64_t *olen, const int64_t *tlen, REAL *gl, int B, int M, int L, int T)
{
    for (int64_t x = 0; x < (int64_t)B * L * T; x++) gl[x] = 0;
    for (int b = 0; b < B; b++) {
        const int O = (int)olen[b], Tn = (int)tlen[b];
        const REAL *a = alpha + (int64_t)b * M * L, *be = beta + (int64_t)b * M * L;
        const REAL *E = links + (int64_t)b * L * T;
        REAL *g = gl + (int64_t)b * L * T;
        const REAL Z = be[0];
        if (FN(is_inf)(Z)) continue;
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = 0; i < O; i++) {
            for (int k = 0; k < T; k++) {
                int n = i + k + 1;
                if (n >= O) break;
                REAL extra = E[(int64_t)i * T + k] - Z;
                ACC acc = 0;
                for (int t = 0; t + 1 < Tn; t++)
                    acc += EXPF(a[(int64_t)t * L + i] + be[(int64_t)(t + 1) * L + n] + extra);
                g[(int64_t)i * T + k] = (REAL)acc * go[b];
            }
        }
    }
}

/* Viterbi (dag_best_alignment.cu:95-116).  `width` is the reference's TRANS_BLOCK_SIZE
 * (config 1..4 -> 4/8/16/32).  Lane x scans delta = x+1, x+1+width, ... keeping the first
 * strict maximum; lanes are merged by a shuffle-down tree with strict '>' so on ties the
 * lane with the smaller bit-reversed index wins; inside a lane the smaller delta wins. */
static inline int FN(bitrev)(int x, int bits)
{
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

void FN(oracle_viterbi)(const REAL *match, const REAL *links, const int64_t *olen, const int64_t *tlen,
                        REAL *alpha, int32_t *trace, int32_t *path, int B, int M, int L, int T, int width)
{
    const REAL NI = (REAL)(-INFINITY);
    int bits = 0;
    while ((1 << bits) < width) bits++;
    for (int64_t x = 0; x < (int64_t)B * M * L; x++) { alpha[x] = NI; trace[x] = 0; }
    for (int64_t x = 0; x < (int64_t)B * L; x++) path[x] = -1;
    for (int b = 0; b < B; b++) {
        const int O = (int)olen[b], Tn = (int)tlen[b];
        const REAL *m = match + (int64_t)b * M * L;
        const REAL *E = links + (int64_t)b * L * T;
        REAL *a = alpha + (int64_t)b * M * L;
        int32_t *tr = trace + (int64_t)b * M * L;
        a[0] = m[0];
        for (int t = 1; t < Tn; t++) {
            const REAL *ap = a + (int64_t)(t - 1) * L;
#pragma omp parallel for schedule(static) if (O - t > 128)
            for (int j = t; j < O; j++) {
                int maxdelta = j < T ? j : T;
                REAL best = NI;
                int bestidx = -1, bestrank = 0, bestd = 0;
                for (int d = 1; d <= maxdelta; d++) {
                    REAL v = ap[j - d] + E[(int64_t)(j - d) * T + (d - 1)];
                    int rank = FN(bitrev)((d - 1) & (width - 1), bits);
                    int better = 0;
                    if (v > best) better = 1;
                    else if (v == best && bestidx >= 0 && (rank < bestrank || (rank == bestrank && d < bestd))) better = 1;
                    if (better) { best = v; bestidx = j - d; bestrank = rank; bestd = d; }
                }
                a[(int64_t)t * L + j] = best + m[(int64_t)t * L + j];
                tr[(int64_t)t * L + j] = bestidx;
            }
        }
        /* backtrace (dag_best_alignment.cu:178-184); an unreachable end cell is a device
         * assert in the reference (:118) -- here the walk simply stops. */
        int pos = O - 1;
        for (int i = Tn - 1; i >= 0 && pos >= 0; i--) {
            path[(int64_t)b * L + pos] = i;
            pos = (i > 0) ? tr[(int64_t)i * L + pos] : -1;
        }
    }
}

/* logsoftmax + gather (logsoftmax_gather.cu:268-308).  idx is addressed through element
 * strides so the stride-0 expanded index tensor of the criterion (nat_dag_loss.py:127)
 * is honoured.  probs (may be NULL) receives softmax probabilities (the in-place
 * overwrite of the logits when a gradient is required). */
void FN(oracle_logsoftmax_gather)(const REAL *logits, const int64_t *idx, int64_t isb, int64_t isl, int64_t iss,
                                  REAL *out, REAL *probs, int B, int L, int V, int S)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)B * L; r++) {
        const REAL *x = logits + r * V;
        int b = (int)(r / L), l = (int)(r % L);
        REAL mx = (REAL)(-INFINITY);
        for (int v = 0; v < V; v++) if (x[v] > mx) mx = x[v];
        ACC s = 0;
        for (int v = 0; v < V; v++) s += EXPF(x[v] - mx);
        REAL ls = LOGF((REAL)s);
        for (int k = 0; k < S; k++) {
            int64_t id = idx[b * isb + l * isl + k * iss];
            out[r * S + k] = (x[id] - mx) - ls;
        }
        if (probs) for (int v = 0; v < V; v++) probs[r * V + v] = EXPF(x[v] - mx) / (REAL)s;
    }
}

/* gather backward (dag_loss.py:293-295): g_in = probs * (-sum_s g) ; g_in[idx[s]] += g[s] */
void FN(oracle_logsoftmax_gather_backward)(const REAL *probs, const int64_t *idx, int64_t isb, int64_t isl, int64_t iss,
                                           const REAL *gout, REAL *gin, int B, int L, int V, int S)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)B * L; r++) {
        int b = (int)(r / L), l = (int)(r % L);
        ACC sum = 0;
        for (int k = 0; k < S; k++) sum += gout[r * S + k];
        REAL neg = (REAL)(-sum);
        for (int v = 0; v < V; v++) gin[r * V + v] = probs[r * V + v] * neg;
        for (int k = 0; k < S; k++) gin[r * V + idx[b * isb + l * isl + k * iss]] += gout[r * S + k];
    }
}

#undef FN
#undef CAT
#undef CAT_
