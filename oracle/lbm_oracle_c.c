/*
 * C restatement of the reference's semantics "C" step (TEST INFRASTRUCTURE ONLY -- see oracle/lbm_oracle.py; the
 * product never links or calls this).  Same two passes, same operation order as the NumPy oracle's step_C, which
 * follows MRT_GPU.py funRT (SRT :338-422, TRT :426-531, MRT :535-662, Smagorinsky :570-589) and funBC (:664-699)
 * evaluated in fp64.  Built with -O2 -ffp-contract=off (no FMA contraction, no fast-math) so that it reproduces the
 * NumPy oracle bit for bit (tests/test_oracle.py); its only purpose is speed (OpenMP over x; every store is owned by one
 * node, so threading does not change a bit): parity runs at 384^2 x 1000 steps and 1024^2 finish in seconds, and it
 * serves as the "port" CPU baseline when oracle/_ref is absent.
 *
 * Arrays use the reference host convention: f[k][x][y] (y fastest), y == 0 is the lid.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const int CX[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};            /* MRT.py:138 */
static const int CY[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const double M_GS[9][9] = {                                   /* MRT_GPU.py:593-601 */
    {1, 1, 1, 1, 1, 1, 1, 1, 1},     {-4, -1, -1, -1, -1, 2, 2, 2, 2}, {4, -2, -2, -2, -2, 1, 1, 1, 1},
    {0, 1, 0, -1, 0, 1, -1, -1, 1},  {0, -2, 0, 2, 0, 1, -1, -1, 1},   {0, 0, 1, 0, -1, 1, 1, -1, -1},
    {0, 0, -2, 0, 2, 1, 1, -1, -1},  {0, 1, -1, 1, -1, 0, 0, 0, 0},    {0, 0, 0, 0, 0, 1, -1, 1, -1}};

enum { SRT = 0, TRT = 1, MRT = 2 };

typedef struct {
    int nx, ny, collision, turb;
    double uLB, omega, omega_e, omega_eps, omega_q, omegam;
} oracle_params;

#define IDX(k, x, y) (((size_t)(k) * nx + (x)) * ny + (y))
#define I2(x, y) ((size_t)(x) * ny + (y))

/* one funRT + funBC; fin, ftemp, feq: [9][nx][ny]; rho: [nx][ny]; u: [2][nx][ny] */
static void step_once(const oracle_params* p, double* fin, double* ftemp, double* feq, double* rho, double* u) {
    const int nx = p->nx, ny = p->ny;
    double T[9], MINV[9][9];
    int k, j;
    T[0] = 4.0 / 9.0;
    for (k = 1; k < 5; ++k) T[k] = 1.0 / 9.0;
    for (k = 5; k < 9; ++k) T[k] = 1.0 / 36.0;
    {   /* MRT_GPU.py:604-612 */
        const double a = 1.0 / 9, b = 1.0 / 36, c = 1.0 / 18, d = 1.0 / 6, e = 1.0 / 12, q = 1.0 / 4;
        const double mi[9][9] = {{a, -a, a, 0, 0, 0, 0, 0, 0},       {a, -b, -c, d, -d, 0, 0, q, 0},  {a, -b, -c, 0, 0, d, -d, -q, 0},
                                 {a, -b, -c, -d, d, 0, 0, q, 0},     {a, -b, -c, 0, 0, -d, d, -q, 0}, {a, c, b, d, e, d, e, 0, q},
                                 {a, c, b, -d, -e, d, e, 0, -q},     {a, c, b, -d, -e, -d, -e, 0, q}, {a, c, b, d, e, -d, -e, 0, -q}};
        memcpy(MINV, mi, sizeof(mi));
    }
    /* every ftemp slot is written by exactly one source node and every other store is per node: the x loop is
     * race-free and order-independent, so OpenMP changes nothing in the result */
#pragma omp parallel for schedule(static) private(k, j)
    for (int x = 0; x < nx; ++x) {
        for (int y = 0; y < ny; ++y) {
            double f[9], fe[9], fpost[9];
            double omega_nu = p->omega;
            for (k = 0; k < 9; ++k) f[k] = fin[IDX(k, x, y)];
            if (p->turb == 1) {                                      /* MRT_GPU.py:570-589, previous-step feq_g / rho_g */
                const double Cs2 = 0.025;
                double product1 = 0.0, product2 = 0.0;
                for (k = 0; k < 9; ++k) {
                    const int cc = CX[k] * CY[k];
                    product1 = cc * f[k] + product1;
                    product2 = cc * feq[IDX(k, x, y)] + product2;
                }
                const double Qmf = product1 - product2;
                const double tau0 = 1.0 / p->omega;
                const double tau = 0.5 * (tau0 + sqrt((tau0 * tau0 + (18 * 1.4142 * Cs2 * fabs(Qmf)) / rho[I2(x, y)])));
                omega_nu = 1.0 / tau;
            }
            double rho_l = f[0];                                     /* :615-619 */
            for (k = 1; k < 9; ++k) rho_l = rho_l + f[k];
            double sx = CX[0] * f[0], sy = CY[0] * f[0];
            for (k = 1; k < 9; ++k) { sx = sx + CX[k] * f[k]; sy = sy + CY[k] * f[k]; }
            double ux = sx / rho_l, uy = sy / rho_l;
            if (x == 0 || x == nx - 1 || y == ny - 1) { ux = 0; uy = 0; }              /* :622-625 */
            if (y == 0) {                                                             /* :626-631 */
                rho_l = f[0] + f[1] + f[3] + 2 * (f[2] + f[5] + f[6]);
                ux = p->uLB; uy = 0;
            }
            rho[I2(x, y)] = rho_l;
            u[IDX(0, x, y)] = ux; u[IDX(1, x, y)] = uy;
            const double usqr = ux * ux + uy * uy;
            for (k = 0; k < 9; ++k) {                                                 /* :649-652 */
                const double cu = (CX[k] * ux + CY[k] * uy);
                fe[k] = rho_l * T[k] * (1. + 3.0 * cu + 9 * 0.5 * cu * cu - 3.0 * 0.5 * usqr);
                feq[IDX(k, x, y)] = fe[k];
            }
            if (p->collision == SRT) {                                                /* :413 */
                for (k = 0; k < 9; ++k) fpost[k] = f[k] - omega_nu * (f[k] - fe[k]);
            } else if (p->collision == TRT) {                                         /* :455-462, 514-527 */
                double fplus[9], fminus[9], feplus[9], feminus[9];
                static const int pa[4] = {2, 5, 6, 1}, po[4] = {4, 7, 8, 3};
                for (j = 0; j < 4; ++j) {
                    const int a = pa[j], o = po[j];
                    fplus[a] = 0.5 * (f[a] + f[o]); fplus[o] = fplus[a];
                    fminus[a] = 0.5 * (f[a] - f[o]); fminus[o] = -fminus[a];
                    feplus[a] = 0.5 * (fe[a] + fe[o]); feplus[o] = feplus[a];
                    feminus[a] = 0.5 * (fe[a] - fe[o]); feminus[o] = -feminus[a];
                }
                fplus[0] = f[0]; fminus[0] = 0; feplus[0] = fe[0]; feminus[0] = 0;
                for (k = 0; k < 9; ++k) fpost[k] = f[k] - omega_nu * (fplus[k] - feplus[k]) - p->omegam * (fminus[k] - feminus[k]);
            } else {                                                                  /* :633-648, 655 */
                double m[9], meq[9];
                for (k = 0; k < 9; ++k) {
                    double acc = M_GS[k][0] * f[0];
                    for (j = 1; j < 9; ++j) acc = acc + M_GS[k][j] * f[j];
                    m[k] = acc;
                }
                const double jx = m[3], jy = m[5];
                meq[0] = rho_l;
                meq[1] = -2.0 * rho_l + 3.0 * (jx * jx + jy * jy);
                meq[2] = -3.0 * (jx * jx + jy * jy) + rho_l + 9.0 * (jx * jx * jy * jy);
                meq[4] = -jx + 3.0 * (jx * jx * jx);
                meq[6] = -jy + 3.0 * (jy * jy * jy);
                meq[7] = jx * jx - jy * jy;
                meq[8] = jx * jy;
                meq[3] = m[3]; meq[5] = m[5];
                const double ov[9] = {0.0, p->omega_e, p->omega_eps, 0.0, p->omega_q, 0.0, p->omega_q, omega_nu, omega_nu};
                for (k = 0; k < 9; ++k) m[k] = m[k] - ov[k] * (m[k] - meq[k]);
                for (k = 0; k < 9; ++k) {
                    double acc = MINV[k][0] * m[0];
                    for (j = 1; j < 9; ++j) acc = acc + MINV[k][j] * m[j];
                    fpost[k] = acc;
                }
            }
            for (k = 0; k < 9; ++k) {                                                 /* push, :654-656 */
                const int xt = x + CX[k], yt = y - CY[k];
                if (xt >= 0 && xt < nx && yt >= 0 && yt < ny) ftemp[IDX(k, xt, yt)] = fpost[k];
            }
        }
    }
    /* funBC :664-699, per node: x-block then y-block, then fin = ftemp */
#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x) {
        for (int y = 0; y < ny; ++y) {
            if (!(x == 0 || x == nx - 1 || y == 0 || y == ny - 1)) continue;
#define FT(k) ftemp[IDX(k, x, y)]
#define FE(k) feq[IDX(k, x, y)]
            if (x == 0) {
                FT(1) = FE(1) - FE(3) + FT(3);
                FT(5) = FE(5) - FE(7) + FT(7);
                FT(8) = FE(8) - FE(6) + FT(6);
            } else if (x == nx - 1) {
                FT(3) = -FE(1) + FE(3) + FT(1);
                FT(6) = -FE(8) + FE(6) + FT(8);
                FT(7) = -FE(5) + FE(7) + FT(5);
            }
            if (y == ny - 1) {
                FT(2) = -FE(4) + FE(2) + FT(4);
                FT(5) = -FE(7) + FE(5) + FT(7);
                FT(6) = -FE(8) + FE(6) + FT(8);
            } else if (y == 0) {
                FT(4) = -FE(2) + FE(4) + FT(2);
                FT(7) = -FE(5) + FE(7) + FT(5);
                FT(8) = -FE(6) + FE(8) + FT(6);
            }
        }
    }
    memcpy(fin, ftemp, sizeof(double) * 9 * (size_t)nx * ny);
}

/* Advance nsteps.  All five arrays are caller-owned and persist between calls (the reference's device arrays). */
void oracle_step_C(const oracle_params* p, double* fin, double* ftemp, double* feq, double* rho, double* u, int nsteps) {
    for (int i = 0; i < nsteps; ++i) step_once(p, fin, ftemp, feq, rho, u);
}
