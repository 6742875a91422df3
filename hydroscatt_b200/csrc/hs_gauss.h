// hs_gauss.h -- host-side Gauss-Legendre table generation (GAUSS).
#pragma once
#include <cmath>
#include <vector>

namespace hs {
// Gauss-Legendre rule of order n on [-1,1] (Newton on P_n; the published GAUSS scheme): nodes ascending, any n >= 1.
inline void gauss_full(int n, double *z, double *w)
{
    int ind = n % 2, k = n / 2 + ind;
    double f = (double)n, x = 0.0;
    for (int i = 1; i <= k; i++) {
        int m = n + 1 - i;
        if (i == 1) x = 1.0 - 2.0 / ((f + 1.0) * f);
        if (i == 2) x = (z[n - 1] - 1.0) * 4.0 + z[n - 1];
        if (i == 3) x = (z[n - 2] - z[n - 1]) * 1.6 + z[n - 2];
        if (i > 3) x = (z[m] - z[m + 1]) * 3.0 + z[m + 2];
        if (i == k && ind == 1) x = 0.0;
        int niter = 0;
        double check = 1e-16, pa = 0, pb, pc;
        for (;;) {
            pb = 1.0;
            niter++;
            if (niter > 100) check *= 10.0;
            pc = x;
            double dj = 1.0;
            for (int j = 2; j <= n; j++) {
                dj += 1.0;
                pa = pb;
                pb = pc;
                pc = x * pb + (x * pb - pa) * (dj - 1.0) / dj;
            }
            pa = 1.0 / ((pb - x * pc) * f);
            pb = pa * pc * (1.0 - x * x);
            x -= pb;
            if (!(fabs(pb) > check * fabs(x))) break;
        }
        z[m - 1] = x;
        w[m - 1] = 2.0 * pa * pa * (1.0 - x * x);
        if (!(i == k && ind == 1)) { z[i - 1] = -z[m - 1]; w[i - 1] = w[m - 1]; }
    }
}

// even order n: the n/2 positive nodes ascending and their weights
inline void gauss_positive(int n, double *xp, double *wp)
{
    std::vector<double> z(n), w(n);
    gauss_full(n, z.data(), w.data());
    int g = n / 2;
    for (int j = 0; j < g; j++) { xp[j] = z[g + j]; wp[j] = w[g + j]; }
}

}  // namespace hs
