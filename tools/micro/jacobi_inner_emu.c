#include <complex.h>
#include <math.h>
/* npass cyclic sweeps of two-sided Jacobi on the Hermitian n x n matrix g (row-major),
   accumulating j (n x n, row-major, identity on entry by the caller).  Returns rotations. */
int inner_sweeps(int n, double complex *g, double complex *j, double tol2, double floor2, int npass)
{
    int rot = 0;
    for (int pass = 0; pass < npass; ++pass) {
        int did = 0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double a = creal(g[p * n + p]), b = creal(g[q * n + q]);
                double complex gpq = g[p * n + q];
                double m2 = creal(gpq) * creal(gpq) + cimag(gpq) * cimag(gpq);
                double big = a > b ? a : b, small = a > b ? b : a;
                if (!(m2 > big * (tol2 * small + floor2))) continue;
                double m = sqrt(m2);
                double complex ph = gpq / m;               /* e^{i phi} */
                double zeta = (b - a) / (2.0 * m);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                /* R = [[c, s],[-s conj(ph), c conj(ph)]] */
                double complex r00 = c, r01 = s, r10 = -s * conj(ph), r11 = c * conj(ph);
                for (int k = 0; k < n; ++k) {               /* columns: G <- G R, J <- J R */
                    double complex x = g[k * n + p], y = g[k * n + q];
                    g[k * n + p] = x * r00 + y * r10;
                    g[k * n + q] = x * r01 + y * r11;
                    x = j[k * n + p]; y = j[k * n + q];
                    j[k * n + p] = x * r00 + y * r10;
                    j[k * n + q] = x * r01 + y * r11;
                }
                for (int k = 0; k < n; ++k) {               /* rows: G <- R^H G */
                    double complex x = g[p * n + k], y = g[q * n + k];
                    g[p * n + k] = conj(r00) * x + conj(r10) * y;
                    g[q * n + k] = conj(r01) * x + conj(r11) * y;
                }
                g[p * n + q] = 0; g[q * n + p] = 0;
                g[p * n + p] = creal(g[p * n + p]); g[q * n + q] = creal(g[q * n + q]);
                ++did;
            }
        rot += did;
        if (!did) break;
    }
    return rot;
}
