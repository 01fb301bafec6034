/*
 * CPU ORACLE (C, OpenMP) -- TEST / BASELINE INFRASTRUCTURE ONLY.  NOT A PRODUCT PATH.
 *
 * Restates the explicit P1DG shallow-water stage of the reference for the CPU
 * baseline that bench.py times next to the GPU path ("cpu_baseline", kind "port")
 * and for parity checks at sizes the numpy oracle (oracle/swe_oracle.py) is too
 * slow for.  Only tests/, __graft_entry__.smoke() and bench.py may load it.
 *
 * Follows (paths relative to /root/reference):
 *   thetis/shallowwater_eq.py:353-393   ExternalPressureGradientTerm (dg branch)
 *   thetis/shallowwater_eq.py:416-450   HUDivTerm (by-parts branch)
 *   thetis/shallowwater_eq.py:470-510   HorizontalAdvectionTerm (+ Lax-Friedrichs)
 *   thetis/shallowwater_eq.py:623-634   CoriolisTerm        :679-701 QuadraticDragTerm (Manning)
 *   thetis/shallowwater_eq.py:734-740   LinearDragTerm      :232-272 get_bnd_functions
 *   thetis/utility.py:975-996           DepthExpression (incl. wetting-drying displacement)
 *   thetis/equation.py:99-105           mass term           thetis/rungekutta.py:929-946 stage update
 *
 * Like the TSFC-generated kernels it replaces, every integral is evaluated by
 * quadrature (6-point degree-3 cell rule, 2-point Gauss on facets) and the 3x3
 * mass system of each cell is solved by elimination.  The loop is fused per
 * cell (gather form), which is already far leaner than the reference's
 * assemble + PETSc solve + vector assigns; it is a generous CPU baseline.
 * Parity status: checked against oracle/swe_oracle.py in tests/test_c_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t n_cells, n_vertices, n_bfacets;
    const double *coords;      /* [nv*2] */
    const int32_t *cells;      /* [nt*3] CCW */
    const int32_t *nbr;        /* [nt*3] >=0 neighbour, <0 -(1+bfacet) */
    const int8_t *nbr_lf;      /* [nt*3] */
    const int32_t *bf_opcode;  /* [nb] OR of 1=elev 2=uv 4=un 8=flux, 0 = closed */
    const double *bf_elev;     /* [nb*2] external elevation at the two facet nodes */
    const double *bf_uv;       /* [nb*4] */
    const double *bf_un;       /* [nb*2] */
    const double *bf_flux;     /* [nb*2] */
    const double *bf_len;      /* [nb] boundary_len[marker] of the facet's marker */
    const double *bath;        /* [nv] */
    const double *coriolis;    /* [nv] or NULL */
    const double *manning;     /* [nv] or NULL */
    double linear_drag;        /* constant, 0 = none */
    double g, lf_sigma, eps2;
    int nonlinear, lf_on;
    int wd_on;                 /* use_wetting_and_drying (utility.py:975-985) */
    double wd_alpha;
    const double *wind;        /* [nv*2] wind stress or NULL (WindStressTerm, shallowwater_eq.py:643-649) */
    double rho0;
} swe_problem;

static const double QL[6][3] = {
    {0.109039009072877, 0.659027622374092, 0.231933368553031},
    {0.231933368553031, 0.659027622374092, 0.109039009072877},
    {0.109039009072877, 0.231933368553031, 0.659027622374092},
    {0.659027622374092, 0.231933368553031, 0.109039009072877},
    {0.231933368553031, 0.109039009072877, 0.659027622374092},
    {0.659027622374092, 0.109039009072877, 0.231933368553031}};
static const double GS[2] = {0.21132486540518711775, 0.78867513459481288225};

/* DepthExpression.get_total_depth, nonlinear case: hl = bathymetry + eta */
static inline double depth_nl(const swe_problem *P, double hl) {
    if (P->wd_on) return hl + 0.5 * (sqrt(hl * hl + P->wd_alpha * P->wd_alpha) - hl);
    return hl;
}

static void solve3(double M[3][3], double b[3]) {
    /* Gaussian elimination without pivoting (SPD mass matrix) */
    for (int k = 0; k < 3; ++k) {
        for (int i = k + 1; i < 3; ++i) {
            double f = M[i][k] / M[k][k];
            for (int j = k; j < 3; ++j) M[i][j] -= f * M[k][j];
            b[i] -= f * b[k];
        }
    }
    for (int i = 2; i >= 0; --i) {
        for (int j = i + 1; j < 3; ++j) b[i] -= M[i][j] * b[j];
        b[i] /= M[i][i];
    }
}

/* out[c*9..] = a0*u0 + a1*u + bdt * M^-1 R(u)   (u0 may be NULL) */
void swe_oracle_stage(const swe_problem *P, double a0, double a1, double bdt, const double *u, const double *u0,
                      double *out) {
    const double g = P->g;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < P->n_cells; ++c) {
        const double *r = u + c * 9;
        double x[3], y[3], b[3], ux[3], uy[3], et[3], f[3] = {0, 0, 0}, mu[3] = {0, 0, 0};
        double twx[3] = {0, 0, 0}, twy[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) {
            const int32_t v = P->cells[c * 3 + a];
            x[a] = P->coords[2 * v];
            y[a] = P->coords[2 * v + 1];
            b[a] = P->bath[v];
            if (P->coriolis) f[a] = P->coriolis[v];
            if (P->manning) mu[a] = P->manning[v];
            if (P->wind) { twx[a] = P->wind[2 * v]; twy[a] = P->wind[2 * v + 1]; }
            ux[a] = r[2 * a];
            uy[a] = r[2 * a + 1];
            et[a] = r[6 + a];
        }
        const double area = 0.5 * ((x[1] - x[0]) * (y[2] - y[0]) - (y[1] - y[0]) * (x[2] - x[0]));
        double gx[3], gy[3], nx[3], ny[3], len[3];
        for (int i = 0; i < 3; ++i) {
            const int p = (i + 1) % 3, q = (i + 2) % 3;
            const double ex = x[q] - x[p], ey = y[q] - y[p];
            len[i] = sqrt(ex * ex + ey * ey);
            nx[i] = ey / len[i];
            ny[i] = -ex / len[i];
            gx[i] = -nx[i] * len[i] / (2.0 * area);
            gy[i] = -ny[i] * len[i] / (2.0 * area);
        }
        double Rx[3] = {0, 0, 0}, Ry[3] = {0, 0, 0}, Re[3] = {0, 0, 0};
        double divu = 0.0;
        for (int a = 0; a < 3; ++a) divu += gx[a] * ux[a] + gy[a] * uy[a];
        /* ---- cell integrals by quadrature */
        for (int k = 0; k < 6; ++k) {
            const double *l = QL[k];
            const double w = area / 6.0;
            const double uq = l[0] * ux[0] + l[1] * ux[1] + l[2] * ux[2];
            const double vq = l[0] * uy[0] + l[1] * uy[1] + l[2] * uy[2];
            const double eq = l[0] * et[0] + l[1] * et[1] + l[2] * et[2];
            const double bq = l[0] * b[0] + l[1] * b[1] + l[2] * b[2];
            const double H = P->nonlinear ? depth_nl(P, bq + eq) : bq;
            double sx = 0.0, sy = 0.0;
            if (P->coriolis) {
                const double fq = l[0] * f[0] + l[1] * f[1] + l[2] * f[2];
                sx += fq * vq;
                sy -= fq * uq;
            }
            if (P->manning) {
                const double m = l[0] * mu[0] + l[1] * mu[1] + l[2] * mu[2];
                const double cd = g * m * m / pow(H, 1.0 / 3.0);
                const double k2 = cd * sqrt(uq * uq + vq * vq + P->eps2) / H;
                sx -= k2 * uq;
                sy -= k2 * vq;
            }
            if (P->linear_drag != 0.0) {
                sx -= P->linear_drag * uq;
                sy -= P->linear_drag * vq;
            }
            if (P->wind) {
                /* +tau / (H rho0) */
                sx += (l[0] * twx[0] + l[1] * twx[1] + l[2] * twx[2]) / (H * P->rho0);
                sy += (l[0] * twy[0] + l[1] * twy[1] + l[2] * twy[2]) / (H * P->rho0);
            }
            for (int a = 0; a < 3; ++a) {
                /* PG: +g*eta*div(psi) ; HUDiv: +grad(phi).(H u) */
                Rx[a] += w * g * eq * gx[a];
                Ry[a] += w * g * eq * gy[a];
                Re[a] += w * H * (gx[a] * uq + gy[a] * vq);
                if (P->nonlinear) {
                    const double cf = gx[a] * uq + gy[a] * vq + l[a] * divu;
                    Rx[a] += w * cf * uq;
                    Ry[a] += w * cf * vq;
                }
                Rx[a] += w * l[a] * sx;
                Ry[a] += w * l[a] * sy;
            }
        }
        /* ---- facets */
        for (int i = 0; i < 3; ++i) {
            const int p = (i + 1) % 3, q = (i + 2) % 3;
            const int32_t nb = P->nbr[c * 3 + i];
            for (int k = 0; k < 2; ++k) {
                const double s = GS[k], wp = 1.0 - s, wq = s, w = 0.5 * len[i];
                const double uK = wp * ux[p] + wq * ux[q], vK = wp * uy[p] + wq * uy[q];
                const double eK = wp * et[p] + wq * et[q], bg = wp * b[p] + wq * b[q];
                const double HK = P->nonlinear ? depth_nl(P, bg + eK) : bg;
                double fx, fy, fe = 0.0;
                if (nb >= 0) {
                    const int j = P->nbr_lf[c * 3 + i];
                    const int np_ = (j + 2) % 3, nq_ = (j + 1) % 3;
                    const double *rn = u + (int64_t)nb * 9;
                    const double uN = wp * rn[2 * np_] + wq * rn[2 * nq_];
                    const double vN = wp * rn[2 * np_ + 1] + wq * rn[2 * nq_ + 1];
                    const double eN = wp * rn[6 + np_] + wq * rn[6 + nq_];
                    const double HN = P->nonlinear ? depth_nl(P, bg + eN) : bg;
                    const double hav = 0.5 * (HK + HN);
                    const double jun = (uK - uN) * nx[i] + (vK - vN) * ny[i];
                    const double head = 0.5 * (eK + eN) + sqrt(hav / g) * jun;
                    fx = g * head * nx[i];
                    fy = g * head * ny[i];
                    const double uax = 0.5 * (uK + uN), uay = 0.5 * (vK + vN);
                    const double sq = sqrt(g / hav);
                    fe = hav * ((uax + sq * (eK - eN) * nx[i]) * nx[i] + (uay + sq * (eK - eN) * ny[i]) * ny[i]);
                    if (P->nonlinear) {
                        const double un = uK * nx[i] + vK * ny[i];
                        fx += uax * un;
                        fy += uay * un;
                        if (P->lf_on) {
                            const double gam = 0.5 * fabs(uax * nx[i] + uay * ny[i]) * P->lf_sigma;
                            fx += gam * (uK - uN);
                            fy += gam * (vK - vN);
                        }
                    }
                } else {
                    const int64_t gb = -(int64_t)nb - 1;
                    const int op = P->bf_opcode[gb];
                    const double un = uK * nx[i] + vK * ny[i];
                    if (op == 0) {
                        const double head = eK + sqrt(HK / g) * un;
                        fx = g * head * nx[i];
                        fy = g * head * ny[i];
                        if (P->nonlinear && P->lf_on) {
                            const double gam = 0.5 * fabs(un) * P->lf_sigma;
                            fx += gam * 2.0 * un * nx[i];
                            fy += gam * 2.0 * un * ny[i];
                        }
                    } else {
                        double eext = eK, uext = uK, vext = vK;
                        const double elev = wp * P->bf_elev[2 * gb] + wq * P->bf_elev[2 * gb + 1];
                        const double bun = wp * P->bf_un[2 * gb] + wq * P->bf_un[2 * gb + 1];
                        const double bfl = wp * P->bf_flux[2 * gb] + wq * P->bf_flux[2 * gb + 1];
                        const double bux = wp * P->bf_uv[4 * gb] + wq * P->bf_uv[4 * gb + 2];
                        const double buy = wp * P->bf_uv[4 * gb + 1] + wq * P->bf_uv[4 * gb + 3];
                        if ((op & 1) && (op & 2)) { eext = elev; uext = bux; vext = buy; }
                        else if ((op & 1) && (op & 4)) { eext = elev; uext = bun * nx[i]; vext = bun * ny[i]; }
                        else if ((op & 1) && (op & 8)) {
                            eext = elev;
                            const double hext = P->nonlinear ? depth_nl(P, bg + eext) : bg;
                            const double sc = bfl / (hext * P->bf_len[gb]);
                            uext = sc * nx[i]; vext = sc * ny[i];
                        } else if (op & 1) { eext = elev; }
                        else if (op & 2) { uext = bux; vext = buy; }
                        else if (op & 4) { uext = bun * nx[i]; vext = bun * ny[i]; }
                        else {
                            const double sc = bfl / (HK * P->bf_len[gb]);
                            uext = sc * nx[i]; vext = sc * ny[i];
                        }
                        const double unj = (uK - uext) * nx[i] + (vK - vext) * ny[i];
                        const double eta_rie = 0.5 * (eK + eext) + sqrt(HK / g) * unj;
                        fx = g * eta_rie * nx[i];
                        fy = g * eta_rie * ny[i];
                        const double Hext = P->nonlinear ? depth_nl(P, bg + eext) : bg;
                        const double hav = 0.5 * (HK + Hext);
                        const double unav = 0.5 * ((uK + uext) * nx[i] + (vK + vext) * ny[i]);
                        const double un_rie = unav + sqrt(g / hav) * (eK - eext);
                        const double eta_rie2 = 0.5 * (eK + eext) + sqrt(hav / g) * unj;
                        const double h_rie = P->nonlinear ? depth_nl(P, bg + eta_rie2) : bg;
                        fe = h_rie * un_rie;
                        if (P->nonlinear) {
                            const double una = unav + sqrt(g / HK) * (eK - eext);
                            fx += una * 0.5 * (uK + uext);
                            fy += una * 0.5 * (vK + vext);
                        }
                    }
                }
                Rx[p] -= w * wp * fx; Ry[p] -= w * wp * fy; Re[p] -= w * wp * fe;
                Rx[q] -= w * wq * fx; Ry[q] -= w * wq * fy; Re[q] -= w * wq * fe;
            }
        }
        /* ---- mass solve + stage update */
        double *rhs[3] = {Rx, Ry, Re};
        for (int comp = 0; comp < 3; ++comp) {
            double M[3][3];
            for (int a = 0; a < 3; ++a)
                for (int bb = 0; bb < 3; ++bb) M[a][bb] = area / 12.0 * (a == bb ? 2.0 : 1.0);
            solve3(M, rhs[comp]);
        }
        double *o = out + c * 9;
        for (int a = 0; a < 3; ++a) {
            double vx = a1 * ux[a] + bdt * Rx[a];
            double vy = a1 * uy[a] + bdt * Ry[a];
            double ve = a1 * et[a] + bdt * Re[a];
            if (u0) {
                vx += a0 * u0[c * 9 + 2 * a];
                vy += a0 * u0[c * 9 + 2 * a + 1];
                ve += a0 * u0[c * 9 + 6 + a];
            }
            o[2 * a] = vx; o[2 * a + 1] = vy; o[6 + a] = ve;
        }
    }
}

/* n SSPRK33 steps (rungekutta.py:342-347 via Shu-Osher form); state updated in place; work = 2 state arrays */
void swe_oracle_ssprk33(const swe_problem *P, double dt, int nsteps, double *state, double *work) {
    const int64_t n = P->n_cells * 9;
    double *B = work, *C = work + n;
    for (int s = 0; s < nsteps; ++s) {
        swe_oracle_stage(P, 0.0, 1.0, dt, state, NULL, B);
        swe_oracle_stage(P, 0.75, 0.25, 0.25 * dt, B, state, C);
        swe_oracle_stage(P, 1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0 * dt, C, state, B);
        memcpy(state, B, sizeof(double) * n);
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to its children: the timed CPU baseline sets its thread count explicitly */
void swe_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int swe_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
