// Displaced-mass update of the elevation for the explicit wetting-drying step (TB_OPT_WD_DISPLACED_MASS).
//
// With wetting-drying the reference's mass term of the elevation is the nonlinear functional
//     F(eta)_a = int_K (eta + f(b + eta)) phi_a dx,     f(H) = (sqrt(H^2 + alpha^2) - H) / 2
// (shallowwater_eq.py:917-920 = Equation.mass_term + BathymetryDisplacementMassTerm :834-850, utility.py:975-985).
// An explicit Shu-Osher stage that advances THIS functional reads
//     F(eta_new) = a0 F(eta_0) + a1 F(eta_i) + beta dt R_eta(u_i),        a0 + a1 = 1,
// and is solved cell by cell for eta_new.
//
// Formulation used here.  eta + f(b + eta) = Ht(b + eta) - b with the total depth Ht(H) = (sqrt(H^2 + alpha^2) + H) / 2
// (utility.py:987-996), and because a0 + a1 = 1 the bathymetry drops out of the stage equation:
//     S(eta_new) = a0 S(eta_0) + a1 S(eta_i) + M k,     S(eta)_a = sum_q w_q Ht(b_q + eta_q) lambda_a(q),
// on the reference cell (weights sum to 1, the area cancels), M = (I + 1 1^T) / 12 and k = beta dt M_K^-1 R_eta
// = lin - a0 eta_0 - a1 eta_i, where lin is the plain-mass update the stage kernel has already formed.  (The degree-3
// cell rule integrates the P1 mass exactly, so this is the same equation as M eta + s(eta) = ... with s built from f;
// written with Ht it has no cancellation: in an almost dry cell eta + f = -b + alpha^2 / (4 |H|) loses every digit of
// the part that depends on eta, Ht = alpha^2 / (2 (r - H)) does not.)
// Damped Newton iteration from lin; the Jacobian sum_q w_q Ht'(H_q) lambda lambda^T, Ht' = (1 + H / r) / 2 > 0, is
// symmetric positive definite: the equation is the stationarity condition of a strictly convex potential, the root is
// unique, and a step that increases the residual (a jump across the kink of Ht from the flat, dry side) is halved.
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

#define TB_WDM_MAX_IT 60

// total depth Ht(H) and its derivative, free of cancellation on both sides of H = 0
static TB_WDM_FN void tb_wdm_depth(double H, double aa, double *ht, double *dht) {
    const double r = sqrt(H * H + aa);
    if (H >= 0.0) {
        *ht = 0.5 * (r + H);
        *dht = r > 0.0 ? 0.5 * (1.0 + H / r) : 0.5;      // r == 0 only for H = alpha = 0: one-sided value
    } else {
        const double d = r - H;                          // > 0
        *ht = 0.5 * aa / d;
        *dht = 0.5 * aa / (d * r);
    }
}

// S(eta)_a = sum_q w_q Ht(b_q + eta_q) lambda_a(q); alpha^2 = a2 (constant) or (sum_a lambda_a al_a)^2 when al != NULL
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
        double ht, dht;
        tb_wdm_depth(H, aa, &ht, &dht);
        const double f = qw[q] * ht;
        s[0] += f * l0;
        s[1] += f * l1;
        s[2] += f * l2;
    }
}

// eta_io: in = lin (plain-mass update), out = eta_new.  eta0 may be NULL (first stage: a0 = 0).  Requires a0 + a1 = 1.
// Returns the number of residual evaluations, or -1 when TB_WDM_MAX_IT of them did not reach the tolerance (e.g. a
// target below what an empty cell holds has no solution); the last accepted iterate is written either way.
static TB_WDM_FN int tb_wd_displaced_update(double *eta_io, double a0, const double *eta0, double a1, const double *etai,
                                            const double *b, const double *al, double a2, const double *qlam,
                                            const double *qw, int nq) {
    double T[3], s[3], k[3];
    const bool use0 = eta0 && a0 != 0.0;
    for (int a = 0; a < 3; ++a) k[a] = eta_io[a] - a1 * etai[a] - (use0 ? a0 * eta0[a] : 0.0);
    const double sk = k[0] + k[1] + k[2];
    for (int a = 0; a < 3; ++a) T[a] = (k[a] + sk) * (1.0 / 12.0);
    if (use0) {
        tb_wdm_s(eta0, b, al, a2, qlam, qw, nq, s);
        for (int a = 0; a < 3; ++a) T[a] += a0 * s[a];
    }
    tb_wdm_s(etai, b, al, a2, qlam, qw, nq, s);
    for (int a = 0; a < 3; ++a) T[a] += a1 * s[a];

    double e[3] = {eta_io[0], eta_io[1], eta_io[2]};
    double base[3] = {e[0], e[1], e[2]}, step[3] = {0.0, 0.0, 0.0};
    double gprev = 1.0e300, prev = 1.0e300;
    int it = 0, ok = 0, halvings = 0;
    for (; it < TB_WDM_MAX_IT; ++it) {
        // residual G = S(e) - T and Jacobian J = sum_q w_q Ht'(H_q) lambda lambda^T (symmetric positive definite)
        double G0 = -T[0], G1 = -T[1], G2 = -T[2];
        double J00 = 0.0, J11 = 0.0, J22 = 0.0, J01 = 0.0, J02 = 0.0, J12 = 0.0;
        for (int q = 0; q < nq; ++q) {
            const double l0 = qlam[3 * q], l1 = qlam[3 * q + 1], l2 = qlam[3 * q + 2];
            const double H = l0 * (b[0] + e[0]) + l1 * (b[1] + e[1]) + l2 * (b[2] + e[2]);
            double aa = a2;
            if (al) {
                const double a = l0 * al[0] + l1 * al[1] + l2 * al[2];
                aa = a * a;
            }
            double ht, dht;
            tb_wdm_depth(H, aa, &ht, &dht);
            ht *= qw[q];
            dht *= qw[q];
            G0 += ht * l0; G1 += ht * l1; G2 += ht * l2;
            J00 += dht * l0 * l0; J11 += dht * l1 * l1; J22 += dht * l2 * l2;
            J01 += dht * l0 * l1; J02 += dht * l0 * l2; J12 += dht * l1 * l2;
        }
        const double gn = fmax(fabs(G0), fmax(fabs(G1), fabs(G2)));
        // rounding level of the residual: S(e) is a sum of positive terms, S = G + T
        const double noise = 4.0e-15 * (fabs(G0 + T[0]) + fabs(G1 + T[1]) + fabs(G2 + T[2]) + fabs(T[0]) + fabs(T[1]) + fabs(T[2]));
        if (gn <= noise && it > 0) {          // the residual cannot get any smaller
            ok = 1;
            ++it;
            break;
        }
        if (gn > gprev && gn > noise && halvings < 40) {
            // reject: back to the last accepted iterate, half of the step that led here
            ++halvings;
            for (int a = 0; a < 3; ++a) {
                step[a] *= 0.5;
                e[a] = base[a] - step[a];
            }
            continue;
        }
        halvings = 0;
        gprev = gn;
        base[0] = e[0]; base[1] = e[1]; base[2] = e[2];
        // step = J^-1 G by the adjugate of the symmetric 3x3 matrix
        const double c00 = J11 * J22 - J12 * J12, c01 = J02 * J12 - J01 * J22, c02 = J01 * J12 - J02 * J11;
        const double c11 = J00 * J22 - J02 * J02, c12 = J01 * J02 - J00 * J12, c22 = J00 * J11 - J01 * J01;
        const double det = J00 * c00 + J01 * c01 + J02 * c02;
        const double id = 1.0 / det;
        step[0] = (c00 * G0 + c01 * G1 + c02 * G2) * id;
        step[1] = (c01 * G0 + c11 * G1 + c12 * G2) * id;
        step[2] = (c02 * G0 + c12 * G1 + c22 * G2) * id;
        const double dm = fmax(fabs(step[0]), fmax(fabs(step[1]), fabs(step[2])));
        // no finite step (non-finite input, or a target no elevation can meet: S > 0 tends to 0 as the cell dries and
        // the iterates run off to minus infinity): give up on the last accepted iterate
        if (!(dm <= 1.0e12) || !(step[0] == step[0]) || !(step[1] == step[1]) || !(step[2] == step[2])) break;
        e[0] -= step[0]; e[1] -= step[1]; e[2] -= step[2];
        const double em = fmax(1.0, fmax(fabs(e[0]), fmax(fabs(e[1]), fabs(e[2]))));
        // converged, or stagnated at the rounding level
        if (dm <= 1.0e-12 * em || (dm >= 0.5 * prev && dm <= 1.0e-8 * em)) {
            ok = 1;
            ++it;
            break;
        }
        prev = dm;
    }
    if (!ok) {          // the last evaluated point may be a rejected one: hand back the last accepted iterate
        e[0] = base[0]; e[1] = base[1]; e[2] = base[2];
    }
    eta_io[0] = e[0]; eta_io[1] = e[1]; eta_io[2] = e[2];
    return ok ? it : -1;
}
