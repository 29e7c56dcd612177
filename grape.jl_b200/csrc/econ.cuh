// econ.cuh -- the economised (Chebyshev-cut) polynomial of exp(-i x): host-side table, device copy.
#pragma once
#include <cmath>
#include <cstring>
#include "common.cuh"

// ---------------------------------------------------------------------------
// Economised exponential for Hermitian generators (the Chebyshev propagator of the reference's tutorials,
// docs/src/tutorial.md:308, 432, re-expressed in the monomial basis the Krylov-form gradient needs).
// exp(-i x) on [-Theta, Theta] is the Chebyshev series J_0(Theta) + 2 sum_k (-i)^k J_k(Theta) T_k(x / Theta); cut at
// degree m its uniform error is 2 sum_{k>m} |J_k(Theta)| ~ 2 (Theta/2)^(m+1) / (m+1)!, a factor 2^m below the Taylor
// remainder, so a step of ||H dt|| = 0.5 needs degree 12 instead of 15..16.  Written in powers of x the cut series is
//     p_m(x) = sum_{j<=m} g[m][j] (-i x)^j / j!,   g[m][j] = j!/Theta^j sum_{k=j,j+2,..<=m} w_k J_k(Theta) |t_kj|  (real, ~ 1),
// (t_kj: coefficient of y^j in T_k, w_0 = 1, w_k = 2): the chains keep generating the TAYLOR terms (the Krylov vectors
// bh_j, ch_j of dense_kry.cuh) and only weigh them with g when summing the new state; the gradient of the polynomial
// propagator is exact with beta(a,b) g[m][a+b+1].  theta[m] = the largest Theta (<= 1) whose error bound is <= 1e-17.
// Only with every operator Hermitian (spectrum of H_n dt inside [-||H_n dt||, ||H_n dt||]) and only on the Krylov-form
// schedule; the block recursion (:taylor, sub-stepped calls, non-Hermitian generators) keeps the Taylor series.
// ---------------------------------------------------------------------------
constexpr int ECON_MAXM = 20;          // >= KRY_MTMAX (dense.cuh) and the classes of small_sym.cuh
constexpr double ECON_TOL = 1e-17;
struct EconTab {
    double theta[ECON_MAXM + 1];
    double g[ECON_MAXM + 1][ECON_MAXM + 1];
    double ones[ECON_MAXM + 1];
};
__constant__ EconTab c_econ;

inline long double econ_besselj(int k, long double x) {   // power series, x <= 1
    long double term = 1.0L;
    for (int i = 1; i <= k; ++i) term *= (x / 2) / i;
    long double sum = term;
    for (int i = 1; i < 60; ++i) {
        term *= -(x / 2) * (x / 2) / ((long double)i * (k + i));
        sum += term;
        if (fabsl(term) < 1e-40L) break;
    }
    return sum;
}
inline long double econ_err(int m, long double th) {
    long double e = 0.0L;
    for (int k = m + 1; k < m + 40; ++k) e += fabsl(econ_besselj(k, th));
    return 2 * e;
}
inline const EconTab& econ_table() {
    static EconTab tab;
    static bool built = false;
    if (built) return tab;
    memset(&tab, 0, sizeof tab);
    // |t_kj| of the Chebyshev polynomials: T_{k+1} = 2 y T_k - T_{k-1}
    static long double T[ECON_MAXM + 1][ECON_MAXM + 1];
    memset(T, 0, sizeof T);
    T[0][0] = 1.0L;
    T[1][1] = 1.0L;
    for (int k = 2; k <= ECON_MAXM; ++k)
        for (int j = 0; j <= k; ++j) T[k][j] = (j ? 2 * T[k - 1][j - 1] : 0.0L) + T[k - 2][j];   // magnitudes add (signs alternate)
    for (int m = 0; m <= ECON_MAXM; ++m) {
        tab.ones[m] = 1.0;
        for (int j = 0; j <= ECON_MAXM; ++j) tab.g[m][j] = 1.0;
        if (m < 2) continue;
        long double lo = 0.0L, hi = 1.0L;
        if (econ_err(m, hi) <= (long double)ECON_TOL) lo = hi;
        else
            for (int it = 0; it < 70; ++it) {
                const long double mid = (lo + hi) / 2;
                if (econ_err(m, mid) <= (long double)ECON_TOL) lo = mid; else hi = mid;
            }
        tab.theta[m] = (double)lo;
        if ((long double)tab.theta[m] > lo) tab.theta[m] = nextafter(tab.theta[m], 0.0);
        const long double th = lo;
        if (th <= 0.0L) continue;
        long double fj = 1.0L, thj = 1.0L;   // j!, Theta^j
        for (int j = 0; j <= m; ++j) {
            if (j) { fj *= j; thj *= th; }
            long double sacc = 0.0L;
            for (int k = j; k <= m; k += 2) sacc += (k ? 2.0L : 1.0L) * econ_besselj(k, th) * T[k][j];
            tab.g[m][j] = (double)(sacc * fj / thj);
        }
    }
    built = true;
    return tab;
}

