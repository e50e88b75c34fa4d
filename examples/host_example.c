/* Whole fit + predict through the C ABI from plain C (route C).
 *
 *   gcc -std=c99 -Iinclude examples/host_example.c -L2d-gp_b200 -lgp2d -Wl,-rpath,$PWD/2d-gp_b200 -lm -o host_example
 *   ./host_example            # needs a CUDA device
 *
 * A 20 x 20 lattice of "drifters" observing a divergence-free flow (u, v) = (-dpsi/dy, dpsi/dx),
 * psi = exp(-r^2 / 8): kriged back onto the observation sites and onto one far-away point. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "gp2d.h"

int main(void) {
    enum { S = 20, N = S * S, M = N + 1 };
    double *X = malloc(sizeof(double) * 2 * N), *y = malloc(sizeof(double) * 2 * N);
    double *Xs = malloc(sizeof(double) * 2 * M), *mean = malloc(sizeof(double) * 2 * M), *var = malloc(sizeof(double) * 2 * M);
    double lml = 0.0, worst = 0.0;
    int i, info;
    if (!X || !y || !Xs || !mean || !var) return 2;
    for (i = 0; i < N; ++i) {
        const double a = 0.5 * (i % S) - 4.75, b = 0.5 * (i / S) - 4.75, psi = exp(-(a * a + b * b) / 8.0);
        X[2 * i] = a; X[2 * i + 1] = b;
        y[i] = b / 4.0 * psi;              /* u = -dpsi/dy : first block belongs to the first coordinate */
        y[N + i] = -a / 4.0 * psi;         /* v =  dpsi/dx */
        Xs[2 * i] = a; Xs[2 * i + 1] = b;
    }
    Xs[2 * N] = 500.0; Xs[2 * N + 1] = 500.0;
    info = gp2d_fit_predict_host(X, N, y, /*l_df*/ 2.0, /*l_cf*/ 2.0, /*ratio*/ 0.9, /*noise*/ 1e-4, /*jitter*/ 0.0,
                                 Xs, M, /*include_noise*/ 0, mean, var, &lml);
    if (info != 0) {
        fprintf(stderr, "gp2d_fit_predict_host: %s\n", gp2d_error_string(info));
        return 1;
    }
    for (i = 0; i < N; ++i) {
        if (fabs(mean[i] - y[i]) > worst) worst = fabs(mean[i] - y[i]);
        if (fabs(mean[M + i] - y[N + i]) > worst) worst = fabs(mean[M + i] - y[N + i]);
    }
    printf("libgp2d %d: LML %.6f, max |mean - obs| at the sites %.2e, far-field mean %.1e variance %.6f (prior %.6f)\n",
           gp2d_version(), lml, worst, mean[N], var[N], 0.9 / 4.0 + 0.1 / 4.0);
    info = (worst < 1e-2 && fabs(var[N] - 0.25) < 1e-9) ? 0 : 1;
    free(X); free(y); free(Xs); free(mean); free(var);
    return info;
}
