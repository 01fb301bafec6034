// Displaced-mass update of the elevation for the explicit wetting-drying step (TB_OPT_WD_DISPLACED_MASS).
//
// With wetting-drying the reference's mass term of the elevation is the nonlinear functional
//     F(eta)_a = int_K (eta + f(b + eta)) phi_a dx,     f(H) = (sqrt(H^2 + alpha^2) - H) / 2
// (shallowwater_eq.py:917-920 = Equation.mass_term + BathymetryDisplacementMassTerm :834-850, utility.py:975-985).
// An explicit Shu-Osher stage that advances THIS functional reads
//     F(eta_new) = a0 F(eta_0) + a1 F(eta_i) + beta dt R_eta(u_i)
// and is solved cell by cell for eta_new.  Written on the reference cell (weights sum to 1; the area cancels):
//     M eta_new + s(eta_new) = T,   M = (I + 1 1^T) / 12,   s(eta)_a = sum_q w_q f(b_q + eta_q) lambda_a(q),
//     T = M lin + a0 s(eta_0) + a1 s(eta_i),
// where lin = a0 eta_0 + a1 eta_i + beta dt M_K^-1 R_eta is the plain-mass update the stage kernel has already formed.
// Newton iteration from lin; the Jacobian sum_q w_q (1 + f'(H_q)) lambda lambda^T is symmetric positive definite
// because 1 + f' = (1 + H / sqrt(H^2 + alpha^2)) / 2 > 0.
//
// Plain C++ (no CUDA intrinsics): the same source is compiled for the device by nvcc and for the host by g++ in
// tests/test_wd_displaced_mass_host.py, which checks it against oracle.SWEOracle.solve_displaced_mass.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TB_WDM_FN __host__ __device__ __noinline__
#else
#define TB_WDM_FN
#endif

#define TB_WDM_MAX_IT 40

// s(eta)_a = sum_q w_q f(b_q + eta_q) lambda_a(q); alpha^2 = a2 (constant) or (sum_a lambda_a al_a)^2 when al != NULL
static TB_WDM_FN void tb_wdm_s(const double *eta, const double *b, const double *al, double a2, const double *qlam,
                               const double *qw, int nq, double *s) {
    s[0] = s[1] = s[2] = 0.0;
    for (int q = 0; q < nq; ++q) {
        const double l0 = qlam[3 * q], l1 = qlam[3 * q + 1], l2 = qlam[3 * q + 2];
        const double H = l0 * (b[0] + eta[0]) + l1 * (b[1] + eta[1]) + l2 * (b[2] + eta[2]);
        double aa = a2;
        if (al) {
            const double a = l0 * al[0] + l1 * al[1] + l2 * al[2];
            aa = a * a;
        }
        const double f = qw[q] * 0.5 * (sqrt(H * H + aa) - H);
        s[0] += f * l0;
        s[1] += f * l1;
        s[2] += f * l2;
    }
}

// eta_io: in = lin (plain-mass update), out = eta_new.  eta0 may be NULL (first stage: a0 = 0).
// Returns the number of Newton iterations, or -1 when TB_WDM_MAX_IT iterations did not reach the tolerance
// (the last iterate is still written).
static TB_WDM_FN int tb_wd_displaced_update(double *eta_io, double a0, const double *eta0, double a1, const double *etai,
                                            const double *b, const double *al, double a2, const double *qlam,
                                            const double *qw, int nq) {
    double T[3], s[3];
    const double sl = eta_io[0] + eta_io[1] + eta_io[2];
    for (int a = 0; a < 3; ++a) T[a] = (eta_io[a] + sl) * (1.0 / 12.0);
    if (eta0 && a0 != 0.0) {
        tb_wdm_s(eta0, b, al, a2, qlam, qw, nq, s);
        for (int a = 0; a < 3; ++a) T[a] += a0 * s[a];
    }
    if (a1 != 0.0) {
        tb_wdm_s(etai, b, al, a2, qlam, qw, nq, s);
        for (int a = 0; a < 3; ++a) T[a] += a1 * s[a];
    }
    double e[3] = {eta_io[0], eta_io[1], eta_io[2]};
    int it = 0, ok = 0;
    double prev = 1.0e300;
    for (; it < TB_WDM_MAX_IT; ++it) {
        // residual G = M e + s(e) - T and Jacobian J = M + sum_q w_q f'(H_q) lambda lambda^T (symmetric)
        const double se = e[0] + e[1] + e[2];
        double G0 = (e[0] + se) * (1.0 / 12.0) - T[0], G1 = (e[1] + se) * (1.0 / 12.0) - T[1],
               G2 = (e[2] + se) * (1.0 / 12.0) - T[2];
        double J00 = 1.0 / 6.0, J11 = 1.0 / 6.0, J22 = 1.0 / 6.0, J01 = 1.0 / 12.0, J02 = 1.0 / 12.0, J12 = 1.0 / 12.0;
        for (int q = 0; q < nq; ++q) {
            const double l0 = qlam[3 * q], l1 = qlam[3 * q + 1], l2 = qlam[3 * q + 2];
            const double H = l0 * (b[0] + e[0]) + l1 * (b[1] + e[1]) + l2 * (b[2] + e[2]);
            double aa = a2;
            if (al) {
                const double a = l0 * al[0] + l1 * al[1] + l2 * al[2];
                aa = a * a;
            }
            const double r = sqrt(H * H + aa);
            const double f = qw[q] * 0.5 * (r - H);
            // f' = (H / r - 1) / 2; for r == 0 (H = alpha = 0) take the one-sided value -1/2
            const double fp = qw[q] * 0.5 * ((r > 0.0 ? H / r : 0.0) - 1.0);
            G0 += f * l0; G1 += f * l1; G2 += f * l2;
            J00 += fp * l0 * l0; J11 += fp * l1 * l1; J22 += fp * l2 * l2;
            J01 += fp * l0 * l1; J02 += fp * l0 * l2; J12 += fp * l1 * l2;
        }
        // d = J^-1 G by the adjugate of the symmetric 3x3 matrix
        const double c00 = J11 * J22 - J12 * J12, c01 = J02 * J12 - J01 * J22, c02 = J01 * J12 - J02 * J11;
        const double c11 = J00 * J22 - J02 * J02, c12 = J01 * J02 - J00 * J12, c22 = J00 * J11 - J01 * J01;
        const double det = J00 * c00 + J01 * c01 + J02 * c02;
        const double id = 1.0 / det;
        const double d0 = (c00 * G0 + c01 * G1 + c02 * G2) * id;
        const double d1 = (c01 * G0 + c11 * G1 + c12 * G2) * id;
        const double d2 = (c02 * G0 + c12 * G1 + c22 * G2) * id;
        e[0] -= d0; e[1] -= d1; e[2] -= d2;
        const double dm = fmax(fabs(d0), fmax(fabs(d1), fabs(d2)));
        const double em = fmax(1.0, fmax(fabs(e[0]), fmax(fabs(e[1]), fabs(e[2]))));
        // converged (also leaves on NaN: the caller sees it in the state), or stagnated at the rounding level of an
        // ill-conditioned, almost dry cell
        if (!(dm > 1.0e-12 * em) || (dm >= 0.5 * prev && dm <= 1.0e-8 * em)) {
            ok = 1;
            ++it;
            break;
        }
        prev = dm;
    }
    eta_io[0] = e[0]; eta_io[1] = e[1]; eta_io[2] = e[2];
    return ok ? it : -1;
}
