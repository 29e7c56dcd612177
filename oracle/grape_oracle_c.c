/*
 * grape_oracle_c.c -- plain-C CPU restatement of the GRAPE.jl ExpProp +
 * GradGenerator gradient path.  TEST / BASELINE INFRASTRUCTURE ONLY: it is the
 * checker and the timed CPU baseline (bench.py `cpu_baseline`, `--impl
 * reference`), never part of the product path.
 *
 * PARITY STATUS: parity unpinned element-wise (see oracle/grape_oracle.py header);
 * this file is validated against oracle/grape_oracle.py, which in turn is pinned
 * by the reference's known-answer tests (tests/test_oracle_known_answers.py).
 *
 * It executes the algorithm the way the reference does:
 *   per trajectory k and step n, a dense Higham scaling-and-squaring Pade
 *   exponential (what Julia's LinearAlgebra.exp does inside ExpProp.prop_step!,
 *   reference src/optimize.jl:732) of the N x N matrix -i H dt going forward, and
 *   of the dense N(L+1) x N(L+1) GradGenerator block matrix going backward
 *   (src/optimize.jl:881, docs/src/background.md:467-477), no sharing between
 *   trajectories; threads over trajectories like `@threadsif` (src/optimize.jl:720, 876).
 *
 * Matrices are row-major here ([i*N + j]); states [k*N + i]; pulses eps[l*NT + n].
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cd;

/* ---- dense helpers ------------------------------------------------------- */
static void mm(int n, const cd* A, const cd* B, cd* C) { /* C = A*B */
    for (int i = 0; i < n; ++i) {
        cd* c = C + (size_t)i * n;
        for (int j = 0; j < n; ++j) c[j] = 0;
        for (int k = 0; k < n; ++k) {
            const cd a = A[(size_t)i * n + k];
            if (a == 0) continue;
            const cd* b = B + (size_t)k * n;
            for (int j = 0; j < n; ++j) c[j] += a * b[j];
        }
    }
}
static double norm1(int n, const cd* A) {
    double m = 0;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += cabs(A[(size_t)i * n + j]);
        if (s > m) m = s;
    }
    return m;
}
/* solve A X = B in place (B <- X), A destroyed; LU with partial pivoting (zgesv) */
static int solve(int n, cd* A, cd* B) {
    for (int c = 0; c < n; ++c) {
        int piv = c;
        double best = cabs(A[(size_t)c * n + c]);
        for (int r = c + 1; r < n; ++r) {
            double v = cabs(A[(size_t)r * n + c]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0) return 1;
        if (piv != c)
            for (int j = 0; j < n; ++j) {
                cd t = A[(size_t)c * n + j]; A[(size_t)c * n + j] = A[(size_t)piv * n + j]; A[(size_t)piv * n + j] = t;
                t = B[(size_t)c * n + j]; B[(size_t)c * n + j] = B[(size_t)piv * n + j]; B[(size_t)piv * n + j] = t;
            }
        const cd inv = 1.0 / A[(size_t)c * n + c];
        for (int r = c + 1; r < n; ++r) {
            const cd f = A[(size_t)r * n + c] * inv;
            if (f == 0) continue;
            for (int j = c; j < n; ++j) A[(size_t)r * n + j] -= f * A[(size_t)c * n + j];
            for (int j = 0; j < n; ++j) B[(size_t)r * n + j] -= f * B[(size_t)c * n + j];
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        for (int j = 0; j < n; ++j) {
            cd s = B[(size_t)r * n + j];
            for (int c = r + 1; c < n; ++c) s -= A[(size_t)r * n + c] * B[(size_t)c * n + j];
            B[(size_t)r * n + j] = s / A[(size_t)r * n + r];
        }
    }
    return 0;
}

/* ---- expm: Higham (2005) scaling & squaring Pade, degrees 3,5,7,9,13 ------ */
/* work: 6*n*n complex */
static void expm_pade(int n, const cd* Ain, cd* E, cd* work) {
    static const double b3[] = {120, 60, 12, 1};
    static const double b5[] = {30240, 15120, 3360, 420, 30, 1};
    static const double b7[] = {17297280, 8648640, 1995840, 277200, 25200, 1512, 56, 1};
    static const double b9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
    static const double b13[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                 129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
                                 40840800., 960960., 16380., 182., 1.};
    static const double theta[] = {1.495585217958292e-2, 2.539398330063230e-1, 9.504178996162932e-1,
                                   2.097847961257068e0, 5.371920351148152e0};
    const size_t nn = (size_t)n * n;
    cd *A = work, *A2 = work + nn, *A4 = work + 2 * nn, *A6 = work + 3 * nn, *U = work + 4 * nn, *V = work + 5 * nn;
    memcpy(A, Ain, nn * sizeof(cd));
    const double nrm = norm1(n, A);
    int s = 0;
    if (nrm > theta[4]) {
        s = (int)ceil(log2(nrm / theta[4]));
        if (s < 0) s = 0;
        const double sc = ldexp(1.0, -s);
        for (size_t q = 0; q < nn; ++q) A[q] *= sc;
    }
    mm(n, A, A, A2);
    if (nrm <= theta[3]) {
        const double* b; int m;
        if (nrm <= theta[0]) { b = b3; m = 3; }
        else if (nrm <= theta[1]) { b = b5; m = 5; }
        else if (nrm <= theta[2]) { b = b7; m = 7; }
        else { b = b9; m = 9; }
        /* U = A * sum_{odd} b_{2j+1} A^{2j},  V = sum_{even} b_{2j} A^{2j} */
        cd* P = A4;      /* running power A^{2j} */
        cd* T = A6;
        for (size_t q = 0; q < nn; ++q) { U[q] = 0; V[q] = 0; P[q] = 0; }
        for (int i = 0; i < n; ++i) { P[(size_t)i * n + i] = 1; }
        for (int j = 0; 2 * j <= m; ++j) {
            for (size_t q = 0; q < nn; ++q) {
                V[q] += b[2 * j] * P[q];
                if (2 * j + 1 <= m) U[q] += b[2 * j + 1] * P[q];
            }
            if (2 * (j + 1) <= m) { mm(n, P, A2, T); memcpy(P, T, nn * sizeof(cd)); }
        }
        mm(n, A, U, T);
        memcpy(U, T, nn * sizeof(cd));
    } else {
        mm(n, A2, A2, A4);
        mm(n, A4, A2, A6);
        const double* b = b13;
        cd* T = E;   /* scratch */
        /* U = A [A6 (b13 A6 + b11 A4 + b9 A2) + b7 A6 + b5 A4 + b3 A2 + b1 I] */
        for (size_t q = 0; q < nn; ++q) T[q] = b[13] * A6[q] + b[11] * A4[q] + b[9] * A2[q];
        mm(n, A6, T, U);
        for (size_t q = 0; q < nn; ++q) U[q] += b[7] * A6[q] + b[5] * A4[q] + b[3] * A2[q];
        for (int i = 0; i < n; ++i) U[(size_t)i * n + i] += b[1];
        mm(n, A, U, T);
        memcpy(U, T, nn * sizeof(cd));
        /* V = A6 (b12 A6 + b10 A4 + b8 A2) + b6 A6 + b4 A4 + b2 A2 + b0 I */
        for (size_t q = 0; q < nn; ++q) T[q] = b[12] * A6[q] + b[10] * A4[q] + b[8] * A2[q];
        mm(n, A6, T, V);
        for (size_t q = 0; q < nn; ++q) V[q] += b[6] * A6[q] + b[4] * A4[q] + b[2] * A2[q];
        for (int i = 0; i < n; ++i) V[(size_t)i * n + i] += b[0];
    }
    /* E = (V - U)^{-1} (V + U) */
    cd* Q = A2;
    for (size_t q = 0; q < nn; ++q) { Q[q] = V[q] - U[q]; E[q] = V[q] + U[q]; }
    solve(n, Q, E);
    for (int t = 0; t < s; ++t) { mm(n, E, E, A); memcpy(E, A, nn * sizeof(cd)); }
}

/* exported for tests */
void grape_oracle_expm(int n, const double* A, double* E) {
    cd* work = (cd*)malloc(6 * (size_t)n * n * sizeof(cd));
    expm_pade(n, (const cd*)A, (cd*)E, work);
    free(work);
}

typedef struct {
    int K, N, L, NT, G, K_global;
    int functional;   /* 0 SM, 1 RE, 2 SS */
    int ja_kind, gb_kind, gb_nD;
    double lambda_a, lambda_b;
    const double* tlist; const int* gen;
    const double* H0;   /* [G][N*N] complex row-major */
    const double* Hc;   /* [G][L][N*N] */
    const double* shape;/* [L][NT] or NULL */
    const double* psi0; const double* tgt; /* [K][N] complex */
    const double* weights; const double* gb_D; /* [nD][N*N] row-major */
    int k_count;        /* trajectories actually processed (bounded sample), <= K */
    int nt_count;       /* time steps actually processed (bounded sample), <= NT  */
    int nthreads;
} oracle_problem;

static void gen_matrix(const oracle_problem* p, const double* eps, int g, int n, cd* H) {
    const size_t nn = (size_t)p->N * p->N;
    memcpy(H, (const cd*)p->H0 + (size_t)g * nn, nn * sizeof(cd));
    for (int l = 0; l < p->L; ++l) {
        double a = eps[(size_t)l * p->NT + n];
        if (p->shape) a *= p->shape[(size_t)l * p->NT + n];
        const cd* Hl = (const cd*)p->Hc + ((size_t)g * p->L + l) * nn;
        for (size_t q = 0; q < nn; ++q) H[q] += a * Hl[q];
    }
}
static void matvec(int n, const cd* A, const cd* x, cd* y) {
    for (int i = 0; i < n; ++i) {
        cd s = 0;
        for (int j = 0; j < n; ++j) s += A[(size_t)i * n + j] * x[j];
        y[i] = s;
    }
}
static double quadform(int n, const cd* D, const cd* x) {
    double s = 0;
    for (int i = 0; i < n; ++i) {
        cd t = 0;
        for (int j = 0; j < n; ++j) t += D[(size_t)i * n + j] * x[j];
        s += creal(conj(x[i]) * t);
    }
    return s;
}

/* evaluate_gradient! for trajectories [0,k_count) and steps [0,nt_count).
 * storage: [k_count][NT+1][N] complex scratch supplied by caller.
 * Outputs: G [L*NT] (only -2 Re sum_k tau_grad + lambda_a grad_J_a), J_parts[3], tau [K]. */
int grape_oracle_eval_fg(const oracle_problem* p, const double* eps, double* storage_,
                         double* Gout, double* J_parts, double* tau_) {
    const int N = p->N, L = p->L, NT = p->NT, Kc = p->k_count, NTc = p->nt_count;
    const int M = N * (L + 1);
    const size_t nn = (size_t)N * N;
    cd* storage = (cd*)storage_;
    cd* tau = (cd*)tau_;
    const double* tl = p->tlist;
    double* jb = (double*)calloc(Kc, sizeof(double));
    const int use_gb = p->gb_kind != 0;
#ifdef _OPENMP
    if (p->nthreads > 0) omp_set_num_threads(p->nthreads);
#endif
    /* ---- forward sweep, src/optimize.jl:720-753 */
#pragma omp parallel
    {
        cd* H = (cd*)malloc(nn * sizeof(cd));
        cd* U = (cd*)malloc(nn * sizeof(cd));
        cd* work = (cd*)malloc(6 * nn * sizeof(cd));
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < Kc; ++k) {
            const int g = p->gen[k];
            cd* st = storage + (size_t)k * (NT + 1) * N;
            memcpy(st, (const cd*)p->psi0 + (size_t)k * N, N * sizeof(cd));
            const cd* D = use_gb ? (const cd*)p->gb_D + (p->gb_nD == 1 ? 0 : (size_t)k * nn) : NULL;
            if (use_gb) jb[k] = quadform(N, D, st) * ((tl[1] - tl[0]) / 2);
            for (int n = 0; n < NTc; ++n) {
                const double dt = tl[n + 1] - tl[n];
                gen_matrix(p, eps, g, n, H);
                for (size_t q = 0; q < nn; ++q) H[q] *= -I * dt;
                expm_pade(N, H, U, work);
                matvec(N, U, st + (size_t)n * N, st + (size_t)(n + 1) * N);
                if (use_gb) {
                    const int ntl = n + 1;
                    const double w = ntl < NT ? 0.5 * (tl[ntl + 1] - tl[ntl - 1]) : (tl[NT] - tl[NT - 1]) / 2;
                    jb[k] += quadform(N, D, st + (size_t)(n + 1) * N) * w;
                }
            }
            cd t = 0;
            const cd* tg = (const cd*)p->tgt + (size_t)k * N;
            for (int i = 0; i < N; ++i) t += conj(tg[i]) * st[(size_t)NTc * N + i];
            tau[k] = t;
        }
        free(H); free(U); free(work);
    }
    /* ---- functional, src/optimize.jl:755-766 */
    const double Kg = p->K_global > 0 ? p->K_global : p->K;
    cd sigma = 0; double s2 = 0, jbsum = 0;
    for (int k = 0; k < Kc; ++k) {
        const double w = p->weights ? p->weights[k] : 1.0;
        sigma += w * tau[k]; s2 += w * creal(tau[k] * conj(tau[k])); jbsum += jb[k];
    }
    if (p->functional == 0) J_parts[0] = 1.0 - creal((sigma / Kg) * conj(sigma / Kg));
    else if (p->functional == 1) J_parts[0] = 1.0 - creal(sigma) / Kg;
    else J_parts[0] = 1.0 - s2 / Kg;
    J_parts[1] = 0; J_parts[2] = use_gb ? p->lambda_b * jbsum : 0;
    if (p->ja_kind == 1) {
        double ja = 0;
        for (int l = 0; l < L; ++l)
            for (int n = 0; n < NT; ++n) ja += eps[(size_t)l * NT + n] * eps[(size_t)l * NT + n] * (tl[n + 1] - tl[n]);
        J_parts[1] = p->lambda_a * ja;
    }
    if (!Gout) { free(jb); return 0; }

    /* ---- backward sweep with the dense GradGenerator, src/optimize.jl:845-911 */
    double* tg_re = (double*)calloc((size_t)Kc * L * NT, sizeof(double));
    int bad = 0;
#pragma omp parallel
    {
        cd* A = (cd*)malloc(nn * sizeof(cd));
        cd* Gm = (cd*)malloc((size_t)M * M * sizeof(cd));
        cd* E = (cd*)malloc((size_t)M * M * sizeof(cd));
        cd* work = (cd*)malloc(6 * (size_t)M * M * sizeof(cd));
        cd* v = (cd*)malloc(M * sizeof(cd));
        cd* v2 = (cd*)malloc(M * sizeof(cd));
        cd* xi = (cd*)malloc(N * sizeof(cd));
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < Kc; ++k) {
            const int g = p->gen[k];
            const double w = p->weights ? p->weights[k] : 1.0;
            const cd* st = storage + (size_t)k * (NT + 1) * N;
            const cd* tg = (const cd*)p->tgt + (size_t)k * N;
            const cd* D = use_gb ? (const cd*)p->gb_D + (p->gb_nD == 1 ? 0 : (size_t)k * nn) : NULL;
            cd c;
            if (p->functional == 0) c = w * sigma / (Kg * Kg);
            else if (p->functional == 1) c = w / (2 * Kg);
            else c = w * tau[k] / Kg;
            cd* x = v + (size_t)L * N;
            for (int i = 0; i < N; ++i) x[i] = c * tg[i];
            if (use_gb && p->lambda_b != 0) {
                matvec(N, D, st + (size_t)NTc * N, xi);
                for (int i = 0; i < N; ++i) x[i] -= (p->lambda_b * (tl[NT] - tl[NT - 1]) / 2) * xi[i];
            }
            double rho = 0;
            for (int i = 0; i < N; ++i) rho += creal(x[i] * conj(x[i]));
            rho = sqrt(rho);
            if (rho < 1e-100) { bad = 1; rho = 1; }
            for (int i = 0; i < N; ++i) x[i] /= rho;
            for (int n = NTc - 1; n >= 0; --n) {
                const double dt = tl[n + 1] - tl[n];
                gen_matrix(p, eps, g, n, A);
                /* G = [[H^dag, 0, mu1^dag],[0, H^dag, mu2^dag],[0,0,H^dag]] ; exp(-i G (-dt)) = exp(+i dt G) */
                memset(Gm, 0, (size_t)M * M * sizeof(cd));
                for (int b = 0; b <= L; ++b)
                    for (int i = 0; i < N; ++i)
                        for (int j = 0; j < N; ++j)
                            Gm[(size_t)(b * N + i) * M + b * N + j] = I * dt * conj(A[(size_t)j * N + i]);
                for (int l = 0; l < L; ++l) {
                    const cd* Hl = (const cd*)p->Hc + ((size_t)g * L + l) * nn;
                    const double s = p->shape ? p->shape[(size_t)l * NT + n] : 1.0;
                    for (int i = 0; i < N; ++i)
                        for (int j = 0; j < N; ++j)
                            Gm[(size_t)(l * N + i) * M + L * N + j] = I * dt * s * conj(Hl[(size_t)j * N + i]);
                }
                expm_pade(M, Gm, E, work);
                for (int i = 0; i < L * N; ++i) v[i] = 0;     /* GradVector / resetgradvec! */
                matvec(M, E, v, v2);
                const cd* pp = st + (size_t)n * N;
                for (int l = 0; l < L; ++l) {
                    cd d = 0;
                    for (int i = 0; i < N; ++i) d += conj(v2[l * N + i]) * pp[i];
                    tg_re[((size_t)k * L + l) * NT + n] = rho * creal(d);
                }
                for (int i = 0; i < N; ++i) x[i] = v2[L * N + i];
                if (use_gb && p->lambda_b != 0 && n > 0) {
                    matvec(N, D, pp, xi);
                    const double f = p->lambda_b * 0.5 * (tl[n + 1] - tl[n - 1]) / rho;
                    for (int i = 0; i < N; ++i) x[i] -= f * xi[i];
                }
            }
        }
        free(A); free(Gm); free(E); free(work); free(v); free(v2); free(xi);
    }
    /* ---- assembly, src/optimize.jl:574-584, 1003-1011 */
    for (int l = 0; l < L; ++l)
        for (int n = 0; n < NT; ++n) {
            double s = 0;
            for (int k = 0; k < Kc; ++k) s += tg_re[((size_t)k * L + l) * NT + n];
            double gval = -2.0 * s;
            if (p->ja_kind == 1) gval += p->lambda_a * 2.0 * eps[(size_t)l * NT + n] * (tl[n + 1] - tl[n]);
            Gout[(size_t)l * NT + n] = gval;
        }
    free(tg_re); free(jb);
    return bad;
}

int grape_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
