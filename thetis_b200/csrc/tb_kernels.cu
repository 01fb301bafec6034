// Hand-written sm_100a kernels for the explicit P1DG shallow-water path.
//
// One CTA = one patch of TB_P SFC-consecutive cells, one thread per cell.
//   * the patch's cell records (9 doubles/cell), its previous-step records (u0)
//     and its static block (vertex columns + local connectivity) arrive in
//     shared memory by TMA bulk copies (cp.async.bulk + mbarrier);
//   * records of the off-patch facet neighbours (the patch halo) are gathered
//     by all threads while the bulk copies are in flight;
//   * each thread evaluates the volume terms and the three facets of its cell
//     (every interior facet is evaluated from both sides: no atomics, results
//     bit-identical regardless of the partition), applies the closed-form
//     P1 mass inverse and the Shu-Osher update, and the patch is written back
//     with one TMA bulk store.
//
// Arithmetic follows thetis/shallowwater_eq.py (line refs at each term).
#include <algorithm>
#include "tb_internal.h"
#include "tb_device.cuh"
#include "tb_wd_mass.cuh"

#ifndef TB_HALO_SPEC
#define TB_HALO_SPEC 6        // halo elements per thread fetched speculatively (covers NH <= 85)
#endif
#ifndef TB_ID_SPEC
#define TB_ID_SPEC 1          // halo ids per thread loaded up front (covers NH <= TB_P)
#endif
#ifndef TB_PREFETCH_DIST
#define TB_PREFETCH_DIST 592   // patches ahead to warm in L2: one wave of 148 SMs x 4 CTAs
#endif
#ifndef TB_GP_UNROLL
#define TB_GP_UNROLL 2        // unroll factor of the two facet Gauss points
#endif
#ifndef TB_QUAD_UNROLL
#define TB_QUAD_UNROLL 6      // unroll factor of the cell-quadrature loop (code size vs. scheduling freedom)
#endif
#define TB_PRAGMA_(x) _Pragma(#x)
#define TB_UNROLL(n) TB_PRAGMA_(unroll n)
#ifndef TB_MINB
#define TB_MINB 4      // resident CTAs per SM the stage kernel is compiled for (register budget)
#endif

__host__ __device__ inline size_t tb_ids_offset(const TbPatchLayout &pl);

// ------------------------------------------------------------------ constants
__constant__ double c_qlam[TB_MAX_QUAD][3];
__constant__ double c_qw[TB_MAX_QUAD];
// The default degree-3 rule (and FIAT's) is ONE symmetric orbit: its six points are the permutations of a single
// barycentric triple (a, b, c), all weights equal.  The specialised stage kernels use that structure (tb_set_quadrature
// verifies it; any other rule is served by the generic kernel):  c_qsym = {a - c, b - c, c, weight}, and at point k
// node tb_quad_ia(k) carries a, node tb_quad_ib(k) carries b.
__constant__ double c_qsym[4];
static bool g_quad_sym = false;
__host__ __device__ constexpr int tb_quad_ia(int k) { return k == 0 ? 1 : k == 1 ? 1 : k == 2 ? 2 : k == 3 ? 0 : k == 4 ? 2 : 0; }
__host__ __device__ constexpr int tb_quad_ib(int k) { return k == 0 ? 2 : k == 1 ? 0 : k == 2 ? 1 : k == 3 ? 1 : k == 4 ? 0 : 2; }

__global__ void test_math_kernel(const double *x, double *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = tb_rsqrt(x[i]);
    out[n + i] = tb_sqrt(x[i]);
    out[2 * n + i] = tb_rcp(x[i]);
    out[3 * n + i] = tb_rcbrt(x[i]);
}
cudaError_t tb_launch_test_math(const double *x, double *out, int n, cudaStream_t s) {
    if (n > 0) test_math_kernel<<<(n + 255) / 256, 256, 0, s>>>(x, out, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ small helpers
// DepthExpression.get_total_depth for the nonlinear case (utility.py:975-996):
// hl = bathymetry + eta at the evaluation point
__device__ __forceinline__ double wd_depth(double hl, int wd_on, double alpha2) {
    if (wd_on) return 0.5 * (hl + tb_sqrt(hl * hl + alpha2));
    return hl;
}

struct BcExt {
    double eta, ux, uy;
};

// External state of an open boundary at one Gauss point
// (ShallowWaterTerm.get_bnd_functions, shallowwater_eq.py:232-272).
// e_in, ux_in, uy_in: interior traces; (nxs, nys) = normal*len; il = 1/len;
// elev/uv/un/flux: boundary data already interpolated to the point.
template <bool NONLIN>
__device__ __forceinline__ BcExt bc_external(int op, double e_in, double ux_in, double uy_in, double bpt, double elev,
                                             double uvx, double uvy, double un, double flux, double bnd_len, double nxs,
                                             double nys, double il, int wd_on, double alpha2) {
    BcExt r;
    const double nx = nxs * il, ny = nys * il;
    if ((op & TB_BC_ELEV) && (op & TB_BC_UV)) {
        r.eta = elev; r.ux = uvx; r.uy = uvy;
    } else if ((op & TB_BC_ELEV) && (op & TB_BC_UN)) {
        r.eta = elev; r.ux = un * nx; r.uy = un * ny;
    } else if ((op & TB_BC_ELEV) && (op & TB_BC_FLUX)) {
        r.eta = elev;
        const double h_ext = NONLIN ? wd_depth(bpt + elev, wd_on, alpha2) : bpt;
        const double s = flux * tb_rcp(h_ext * bnd_len);
        r.ux = s * nx; r.uy = s * ny;
    } else if (op & TB_BC_ELEV) {
        r.eta = elev; r.ux = ux_in; r.uy = uy_in;
    } else if (op & TB_BC_UV) {
        r.eta = e_in; r.ux = uvx; r.uy = uvy;
    } else if (op & TB_BC_UN) {
        r.eta = e_in; r.ux = un * nx; r.uy = un * ny;
    } else {  // TB_BC_FLUX
        r.eta = e_in;
        const double h_ext = NONLIN ? wd_depth(bpt + e_in, wd_on, alpha2) : bpt;
        const double s = flux * tb_rcp(h_ext * bnd_len);
        r.ux = s * nx; r.uy = s * ny;
    }
    return r;
}

__device__ __forceinline__ double coef_at(const TbCoef &c, const double *cols, int NV, int v, int comp = 0) {
    if (c.mode == 2) return cols[(size_t)(c.col + comp) * NV + v];
    return comp == 0 ? c.v0 : c.v1;
}

// ------------------------------------------------------------------ SWE stage kernel
// Compile-time specialisations (warp-uniform runtime flags cost issue slots and registers in a kernel whose fp64
// pipe and issue port are co-critical):
//   SPEC 0  generic: every optional term behind a runtime flag
//   SPEC 1  no optional cell terms (BASELINE configs 1 and 2)
//   SPEC 2  Manning drag + Coriolis, 6-point cell rule (config 5 without wetting-drying)
//   SPEC 3  SPEC 2 + wetting-drying (config 5)
//   SPEC 4  Coriolis + wind stress + linear drag, 6-point cell rule (config 3, stommel2d; linear equations)
//   SPEC 5  SPEC 2 + horizontal viscosity (tidal set-ups that prescribe a viscosity / sponge, e.g. examples/north_sea)
//   SPEC 6  SPEC 3 + horizontal viscosity
template <int SPEC>
struct StageSpec {
    static constexpr bool generic = SPEC == 0;
    static constexpr bool wd = SPEC == 3 || SPEC == 6;
    static constexpr bool man = SPEC == 2 || SPEC == 3 || SPEC == 5 || SPEC == 6;
    static constexpr bool cor = SPEC >= 2;
    static constexpr bool wind = SPEC == 4;
    static constexpr bool lin = SPEC == 4;
    static constexpr bool visc = SPEC == 5 || SPEC == 6;
};

// grad(u)_ij = d u_i / d x_j of a P1 cell and twice its area, from its record r = [u0x u0y u1x u1y u2x u2y ...] and its
// vertex coordinates: g = {Gxx, Gxy, Gyx, Gyy, 2A}.  ONE routine, written in explicit fma form, serves the cell itself
// and every neighbour (in-patch or halo), so the value a cell sees for a neighbour's gradient does not depend on which
// thread produced it (partition-independent results).
__device__ __forceinline__ void tb_cell_gradient(const double *r, const double *x, const double *y, double *g) {
    const double twoA = fma(x[1] - x[0], y[2] - y[0], -((y[1] - y[0]) * (x[2] - x[0])));
    double gxx = 0.0, gxy = 0.0, gyx = 0.0, gyy = 0.0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const double mx = y[(b + 2) % 3] - y[(b + 1) % 3], my = x[(b + 1) % 3] - x[(b + 2) % 3];   // |e_b| n_b
        gxx = fma(r[2 * b], mx, gxx); gxy = fma(r[2 * b], my, gxy);
        gyx = fma(r[2 * b + 1], mx, gyx); gyy = fma(r[2 * b + 1], my, gyy);
    }
    const double s = -tb_rcp(twoA);      // grad(phi_b) = -N_b / 2A
    g[0] = gxx * s; g[1] = gxy * s; g[2] = gyx * s; g[3] = gyy * s; g[4] = twoA;
}

// Open-boundary fluxes at one Gauss point (shallowwater_eq.py:370-375, 431-442, 498-509).  Rare (only facets of
// open markers), kept out of line so that it does not bloat the hot instruction stream.
template <bool NONLIN>
__device__ __noinline__ void open_boundary_flux(const TbBcTable *bc, int gb, int slot, double wp_, double wq_, double uKx,
                                                double uKy, double eK, double bg, double HK, double nxs, double nys,
                                                double il, double len, double g, int wd_on, double a2, int adv_on,
                                                double *out) {
    const TbBcSlot &bs = bc->slots[slot];
    const int op = bs.opcode;
    const int row = bs.arr_mask ? __ldg(bc->bf_row + gb) : 0;
    double elev = bs.elev, uvx = bs.uvx, uvy = bs.uvy, un = bs.un, flux = bs.flux;
    if (bs.arr_mask & TB_BC_ELEV) elev = wp_ * __ldg(bc->ext_elev + 2 * row) + wq_ * __ldg(bc->ext_elev + 2 * row + 1);
    if (bs.arr_mask & TB_BC_UV) {
        uvx = wp_ * __ldg(bc->ext_uv + 4 * row) + wq_ * __ldg(bc->ext_uv + 4 * row + 2);
        uvy = wp_ * __ldg(bc->ext_uv + 4 * row + 1) + wq_ * __ldg(bc->ext_uv + 4 * row + 3);
    }
    if (bs.arr_mask & TB_BC_UN) un = wp_ * __ldg(bc->ext_un + 2 * row) + wq_ * __ldg(bc->ext_un + 2 * row + 1);
    if (bs.arr_mask & TB_BC_FLUX) flux = wp_ * __ldg(bc->ext_flux + 2 * row) + wq_ * __ldg(bc->ext_flux + 2 * row + 1);
    const BcExt ex = bc_external<NONLIN>(op, eK, uKx, uKy, bg, elev, uvx, uvy, un, flux, bs.bnd_len, nxs, nys, il, wd_on, a2);
    const double ig = tb_rcp(g);
    // PG (:370-375)
    const double dun = (uKx - ex.ux) * nxs + (uKy - ex.uy) * nys;   // un_jump*len
    const double cK = tb_sqrt(g * HK);
    const double t = 0.5 * g * (eK + ex.eta) + cK * dun * il;
    double fx = t * nxs, fy = t * nys;
    // HUDiv (:431-442)
    const double Hext = NONLIN ? wd_depth(bg + ex.eta, wd_on, a2) : bg;
    const double hav = 0.5 * (HK + Hext);
    const double cav = tb_sqrt(g * hav);
    const double usx = uKx + ex.ux, usy = uKy + ex.uy;
    const double usN = usx * nxs + usy * nys;
    const double ejump = eK - ex.eta;
    const double un_rie_len = 0.5 * usN + (cav * tb_rcp(hav)) * ejump * len;
    const double eta_rie = 0.5 * (eK + ex.eta) + (cav * ig) * dun * il;
    const double h_rie = NONLIN ? wd_depth(bg + eta_rie, wd_on, a2) : bg;
    const double fe = h_rie * un_rie_len;
    if (NONLIN && adv_on) {
        // advection (:498-509)
        const double un_a_len = 0.5 * usN + (cK * tb_rcp(HK)) * ejump * len;
        fx += 0.5 * usx * un_a_len;
        fy += 0.5 * usy * un_a_len;
    }
    out[0] = fx;
    out[1] = fy;
    out[2] = fe;
    out[3] = ex.ux;      // external velocity and the 'un' datum: Dirichlet terms of the viscosity (:592-609)
    out[4] = ex.uy;
    out[5] = un;
}

// Quadratic friction of the tangential velocity on a boundary facet (BoundaryDragTerm, shallowwater_eq.py:704-726):
//   f += C_D |u_t| (psi . u_t) ds,  u_t = u - (u.n) n,  2-point Gauss rule like every other facet integral.
// F[0..3] = the facet's momentum flux tested against phi_p (x, y) and phi_q (x, y).  Rare: out of line and after the
// Gauss-point loop of the other boundary terms (nothing extra stays live across it), generic kernel only.
// |u_t| may be exactly zero: library sqrt.
__device__ __noinline__ void boundary_drag_facet(double cd, double upx, double upy, double uqx, double uqy, double nxs,
                                                 double nys, double il, double len, double *F) {
    const double nx = nxs * il, ny = nys * il;
#pragma unroll 1
    for (int gp = 0; gp < 2; ++gp) {
        const double wq_ = gp ? TB_XI2 : TB_XI1, wp_ = 1.0 - wq_;
        const double uKx = wp_ * upx + wq_ * uqx, uKy = wp_ * upy + wq_ * uqy;
        const double un = uKx * nx + uKy * ny;
        const double utx = uKx - un * nx, uty = uKy - un * ny;
        const double m = 0.5 * cd * sqrt(utx * utx + uty * uty) * len;
        F[0] += wp_ * m * utx; F[1] += wp_ * m * uty;
        F[2] += wq_ * m * utx; F[3] += wq_ * m * uty;
    }
}

template <bool NONLIN, int SPEC>
__global__ void __launch_bounds__(TB_P, TB_MINB) swe_stage_kernel(const __grid_constant__ TbSweParams prm) {
    typedef StageSpec<SPEC> SP;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    double *S = reinterpret_cast<double *>(smem + 16);              // [(TB_P+NH)][9] stage state (own + halo)
    double *O = S + (size_t)(TB_P + prm.pl.NH) * 9;                 // [TB_P][9] u0 in, result out
    unsigned char *blk = reinterpret_cast<unsigned char *>(O + TB_P * 9);

    const int tid = threadIdx.x;
    const int patch = prm.patch_list ? __ldg(prm.patch_list + blockIdx.x) : prm.patch_first + (int)blockIdx.x;
    const long long cell0 = (long long)patch * TB_P;
    const int NV = prm.pl.NV;

    // Distributed run, partition-boundary patch: the ghost records this patch reads were stored by the peers' previous
    // stage launch; wait until every peer has published them (TbHaloFused).  The boundary patches are the first CTAs
    // of the launch and their peers pushed early in THEIR previous launch, so the wait is normally already satisfied.
    const bool bpatch = prm.halo != nullptr && (int)blockIdx.x < prm.n_bpatch;
    const unsigned long long epoch = bpatch ? tb_fused_wait(prm.halo) : 0ull;

    // Halo ids of this patch: one coalesced load per thread (the row was prefetched to L2 one wave earlier), staged in
    // shared memory so that the gather below indexes them with cheap 32-bit shared loads instead of one global load
    // plus 64-bit address arithmetic per gathered element (that index arithmetic was 7 % of all executed instructions).
    const int *hid = prm.pl.halo_ids + (long long)patch * prm.pl.NH;
    int *ids_s = reinterpret_cast<int *>(smem + tb_ids_offset(prm.pl));
    int myid[TB_ID_SPEC];
#pragma unroll
    for (int j = 0; j < TB_ID_SPEC; ++j) {
        const int h = j * TB_P + tid;
        myid[j] = (h < prm.pl.NH) ? __ldg(hid + h) : 0;
    }
    const int nh9 = __ldg(prm.pl.halo_cnt + patch) * 9;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        const uint32_t sb = (uint32_t)prm.pl.stride;
        const uint32_t rec = TB_P * 9 * sizeof(double);
        mbar_expect_tx(bar, rec + sb + (prm.u0 ? rec : 0u));
        bulk_g2s(S, prm.u_in + cell0 * 9, rec, bar);
        bulk_g2s(blk, prm.pl.sblk + (long long)patch * prm.pl.stride, sb, bar);
        if (prm.u0) bulk_g2s(O, prm.u0 + cell0 * 9, rec, bar);
    } else if (tid == 32) {
        // warm L2 for the patch that will run on this SM slot one wave later
        const int pb = (int)blockIdx.x + TB_PREFETCH_DIST;
        if (pb < (int)gridDim.x) {
            const int pf = prm.patch_list ? __ldg(prm.patch_list + pb) : prm.patch_first + pb;
            const uint32_t rec = TB_P * 9 * sizeof(double);
            bulk_prefetch_l2(prm.u_in + (long long)pf * TB_P * 9, rec);
            bulk_prefetch_l2(prm.pl.sblk + (long long)pf * prm.pl.stride, (uint32_t)prm.pl.stride);
            if (prm.u0) bulk_prefetch_l2(prm.u0 + (long long)pf * TB_P * 9, rec);
        }
    } else if (tid >= 33 && tid < 36) {
        // ... and its halo-id row (the first dependent load of that CTA's prologue)
        const int pb = (int)blockIdx.x + TB_PREFETCH_DIST;
        if (pb < (int)gridDim.x) {
            const int pf = prm.patch_list ? __ldg(prm.patch_list + pb) : prm.patch_first + pb;
            const int off = (tid - 33) * 32;
            if (off < prm.pl.NH) prefetch_l2(prm.pl.halo_ids + (long long)pf * prm.pl.NH + off);
            if (tid == 33) prefetch_l2(prm.pl.halo_cnt + pf);
        }
    }
    // patch halo: records of off-patch facet neighbours, copied asynchronously (LDGSTS) as soon as the ids are staged
    // -- while the bulk copies are still in flight -- and only waited for in front of the facet loop, so their latency
    // hides behind the bulk-copy wait and the volume terms.
    // Element i of the halo block is double (i % 9) of halo cell (i / 9): S[TB_P*9 + i].
#pragma unroll
    for (int j = 0; j < TB_ID_SPEC; ++j) {
        const int h = j * TB_P + tid;
        if (h < prm.pl.NH) ids_s[h] = myid[j];
    }
    for (int h = TB_ID_SPEC * TB_P + tid; h < prm.pl.NH; h += TB_P) ids_s[h] = __ldg(hid + h);      // very large halos only
    __syncthreads();          // ids staged; mbarrier initialised (thread 0) before anybody waits on it
    {
        double *H = S + TB_P * 9;
#pragma unroll
        for (int j = 0; j < TB_HALO_SPEC; ++j) {
            const unsigned i = (unsigned)(j * TB_P + tid);
            const unsigned h = i / 9u;
            if ((int)i < nh9) cp_async8(H + i, prm.u_in + ((long long)ids_s[h] * 9 + (int)(i - h * 9u)));
        }
        for (int i = TB_HALO_SPEC * TB_P + tid; i < nh9; i += TB_P) {      // very large halos only
            const int h = i / 9;
            cp_async8(H + i, prm.u_in + ((long long)ids_s[h] * 9 + (i - h * 9)));
        }
        cp_async_commit();
    }
    mbar_wait(bar, 0);        // every thread observes the TMA completion itself

    // threads past the last owned cell of the last patch evaluate cell 0 of the patch again (uniform control flow:
    // there is a block-wide barrier in the middle) and their result is discarded
    const bool active = (cell0 + tid) < prm.n_owned;
    const int ct = active ? tid : 0;
    double res[9];
    double ia[4] = {0, 0, 0, 0};      // this cell's share of the fused diagnostics (prm.partials)

    {
        const double *cols = reinterpret_cast<const double *>(blk);
        const uint16_t *cv = reinterpret_cast<const uint16_t *>(blk + prm.pl.off_cv) + ct * 3;
        const int *cn = reinterpret_cast<const int *>(blk + prm.pl.off_cn) + ct * 3;
        const double *my = S + ct * 9;
        const double g = prm.g;
        const int wd_on = SP::generic ? prm.wd_on : (SP::wd ? 1 : 0);
        const bool lf_on = SP::generic ? (prm.lf_on != 0) : true;
        const bool has_cor = SP::generic ? (prm.cor.mode != 0) : SP::cor;
        const bool has_man = SP::generic ? (prm.man.mode != 0) : SP::man;
        const bool has_cd = SP::generic ? (prm.cd.mode != 0) : false;
        const bool has_nik = SP::generic ? (prm.nik.mode != 0) : false;      // nikuradse_bed_roughness (:689-697)
        const bool has_lin = SP::generic ? (prm.lin.mode != 0) : SP::lin;
        const bool has_wind = SP::generic ? (prm.wind.mode != 0) : SP::wind;
        const bool has_pa = SP::generic ? (prm.pa.mode >= 2) : false;
        const bool has_msrc = SP::generic ? (prm.msrc.mode != 0) : false;
        const bool has_vsrc = SP::generic ? (prm.vsrc.mode != 0) : false;
        // ModeSplit2DEquations (shallowwater_eq.py:931-966) has no HorizontalAdvectionTerm although the depth is nonlinear
        const bool adv_on = SP::generic ? (prm.adv_on != 0) : true;
        const bool has_visc = SP::generic ? (prm.visc.mode != 0) : SP::visc;
        const bool graddiv = (SP::generic || SP::visc) ? (prm.graddiv != 0) : false;
        const bool use_quad = SP::generic ? (prm.use_quad != 0) : (SP::man || SP::wd || SP::wind);
        const double a2 = prm.wd_alpha2;

        double ux[3], uy[3], et[3], x[3], y[3], b[3];
        int v[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            ux[a] = my[2 * a];
            uy[a] = my[2 * a + 1];
            et[a] = my[6 + a];
            v[a] = cv[a];
            x[a] = cols[v[a]];
            y[a] = cols[NV + v[a]];
            b[a] = cols[2 * NV + v[a]];
        }
        // scaled outward normals N_i = |e_i| n_i of facet i (opposite vertex i); A*grad(phi_i) = -N_i/2
        double Nx[3], Ny[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int p = (i + 1) % 3, q = (i + 2) % 3;
            Nx[i] = y[q] - y[p];
            Ny[i] = x[p] - x[q];
        }
        const double twoA = (x[1] - x[0]) * (y[2] - y[0]) - (y[1] - y[0]) * (x[2] - x[0]);

        double Rux[3] = {0, 0, 0}, Ruy[3] = {0, 0, 0}, Re[3] = {0, 0, 0};

        // coefficient at local node a: Constant, P1 vertex column of the static block, or (generic kernel only) a
        // genuinely discontinuous P1DG field stored per cell node (tb_set_field_cell)
        auto cf = [&](const TbCoef &c, int a, int comp = 0) -> double {
            if (SP::generic && c.mode == 3) return __ldg(c.cell + ((cell0 + ct) * 3 + a) * c.nc + comp);
            return coef_at(c, cols, NV, v[a], comp);
        };
        // wetting_and_drying_alpha as a P1 field (solver2d.py:279-287): alpha^2 at a point from the nodal values
        const bool var_al = SP::generic && wd_on && prm.wda.mode == 2;
        double al[3] = {0, 0, 0};
        if (var_al) {
#pragma unroll
            for (int a = 0; a < 3; ++a) al[a] = coef_at(prm.wda, cols, NV, v[a]);
        }

        // ---------------- volume terms (closed-form P1 integrals) ----------------
        const double sux = ux[0] + ux[1] + ux[2], suy = uy[0] + uy[1] + uy[2];
        const double se = et[0] + et[1] + et[2];
        double wx[3], wy[3];      // u_b + sum_c u_c
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            wx[a] = ux[a] + sux;
            wy[a] = uy[a] + suy;
        }
        {
            // ExternalPressureGradientTerm cell part (shallowwater_eq.py:361): +g*eta*div(psi)
            const double c = (g * (-1.0 / 6.0)) * se;   // g * (se/3) * (-N/2)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                Rux[a] += c * Nx[a];
                Ruy[a] += c * Ny[a];
            }
        }
        if (!(NONLIN && wd_on)) {
            // HUDivTerm cell part (:422): +grad(phi).(H u);  int H u = A/12 sum_b H_b (u_b + su)
            double Wx = 0, Wy = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double H = NONLIN ? (b[a] + et[a]) : b[a];
                Wx += H * wx[a];
                Wy += H * wy[a];
            }
            Wx *= (-1.0 / 24.0);
            Wy *= (-1.0 / 24.0);
#pragma unroll
            for (int a = 0; a < 3; ++a) Re[a] = fma(Nx[a], Wx, fma(Ny[a], Wy, Re[a]));
        }
        if (NONLIN && adv_on) {
            // HorizontalAdvectionTerm cell part (:478): +div(outer(psi,u)).u
            //   int (grad phi_a . u) u_i = 1/12 sum_j (A grad phi_a)_j T_ji,  T_ji = sum_b u_b,j (u_b,i + su_i)
            //   int phi_a (div u) u_i   = D/12 (u_a,i + su_i),                D = sum_b (A grad phi_b).u_b
            double Txx = 0, Txy = 0, Tyx = 0, Tyy = 0, D = 0;
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
                Txx = fma(ux[bb], wx[bb], Txx);
                Txy = fma(ux[bb], wy[bb], Txy);
                Tyx = fma(uy[bb], wx[bb], Tyx);
                Tyy = fma(uy[bb], wy[bb], Tyy);
                D = fma(Nx[bb], ux[bb], fma(Ny[bb], uy[bb], D));
            }
            // the common factor (-1/2) * (1/12) once on T and D instead of once per node
            Txx *= (-1.0 / 24.0); Txy *= (-1.0 / 24.0); Tyx *= (-1.0 / 24.0); Tyy *= (-1.0 / 24.0);
            D *= (-1.0 / 24.0);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                Rux[a] = fma(Nx[a], Txx, fma(Ny[a], Tyx, fma(D, wx[a], Rux[a])));
                Ruy[a] = fma(Nx[a], Txy, fma(Ny[a], Tyy, fma(D, wy[a], Ruy[a])));
            }
        }
        const double A = 0.5 * twoA;
        if (has_cor) {
            // CoriolisTerm (:632-633): R_x += int f u_y phi, R_y -= int f u_x phi
            double f[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) f[a] = cf(prm.cor, a);
            const double F = f[0] + f[1] + f[2];
            const double fux_ = f[0] * ux[0] + f[1] * ux[1] + f[2] * ux[2];
            const double fuy_ = f[0] * uy[0] + f[1] * uy[1] + f[2] * uy[2];
            const double c = A * (1.0 / 60.0);
            const double bx = F * sux + fux_, by = F * suy + fuy_;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                // int f w phi_a = A/60 [F W + f_a W + sum f_b w_b + F w_a + 2 f_a w_a] = A/60 [bx + f_a (W + w_a) + F w_a + f_a w_a]
                const double ix = bx + f[a] * (wx[a] + ux[a]) + F * ux[a];
                const double iy = by + f[a] * (wy[a] + uy[a]) + F * uy[a];
                Rux[a] = fma(c, iy, Rux[a]);
                Ruy[a] = fma(-c, ix, Ruy[a]);
            }
        }
        if (has_lin) {
            // LinearDragTerm (:734-740): -C u
            double f[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) f[a] = cf(prm.lin, a);
            const double F = f[0] + f[1] + f[2];
            const double fux_ = f[0] * ux[0] + f[1] * ux[1] + f[2] * ux[2];
            const double fuy_ = f[0] * uy[0] + f[1] * uy[1] + f[2] * uy[2];
            const double c = A * (1.0 / 60.0);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                Rux[a] -= c * (F * sux + f[a] * sux + fux_ + F * ux[a] + 2.0 * f[a] * ux[a]);
                Ruy[a] -= c * (F * suy + f[a] * suy + fuy_ + F * uy[a] + 2.0 * f[a] * uy[a]);
            }
        }
        if (has_pa) {
            // AtmosphericPressureTerm (:658-663): -grad(p_a)/rho0;  A*grad p = -1/2 sum_a p_a N_a
            double gx = 0, gy = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double pa = cf(prm.pa, a);
                gx += pa * Nx[a];
                gy += pa * Ny[a];
            }
            const double c = (1.0 / 6.0) * tb_rcp(prm.rho0);   // -( -1/2 ) * (1/3)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                Rux[a] += c * gx;
                Ruy[a] += c * gy;
            }
        }
        if (has_msrc) {
            // MomentumSourceTerm (:805-811)
            double fx[3], fy[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                fx[a] = cf(prm.msrc, a, 0);
                fy[a] = cf(prm.msrc, a, 1);
            }
            const double sx = fx[0] + fx[1] + fx[2], sy = fy[0] + fy[1] + fy[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                Rux[a] += A * (1.0 / 12.0) * (fx[a] + sx);
                Ruy[a] += A * (1.0 / 12.0) * (fy[a] + sy);
            }
        }
        if (has_vsrc) {
            // ContinuitySourceTerm (:824-831)
            double f[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) f[a] = cf(prm.vsrc, a);
            const double s = f[0] + f[1] + f[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) Re[a] += A * (1.0 / 12.0) * (f[a] + s);
        }
        // HorizontalViscosityTerm (shallowwater_eq.py:554-616), symmetric interior penalty.  S = stress / nu:
        // grad(u) or 2 sym(grad(u)), constant per cell.
        const double itA = has_visc ? tb_rcp(twoA) : 0.0;
        // nu at a cell node and the cell's stress / nu are re-read from shared memory in the facet part instead of being
        // carried in registers across the whole kernel (the stage kernel is at its 128-register limit)
        auto nu_at = [&](int a) { return coef_at(prm.visc, cols, NV, cv[a]); };
        // [(TB_P + NH)][5] gradients + 2A of the patch and halo cells, behind the reduction scratch (only allocated
        // when a SIPG coefficient is set: pl.off_hcv >= 0)
        double *Gs = reinterpret_cast<double *>(blk + prm.pl.stride) + (TB_P / 32) * 4;
        // grad-depth viscosity source evaluated inside the Manning cell-rule loop (shares the depth and H^(-1/3))
        const bool gd_merge = has_visc && prm.graddepth != 0 && use_quad && has_man;
        double gdVx = 0.0, gdVy = 0.0, gdnu[3] = {0, 0, 0};
        if (has_visc) {
            double nuv[3], Sxx, Sxy, Syx, Syy;
#pragma unroll
            for (int a = 0; a < 3; ++a) nuv[a] = nu_at(a);
            // grad(u) and 2A of this cell, published for the threads of the facet neighbours
            double gk[5];
            tb_cell_gradient(my, x, y, gk);
            if (active) {
#pragma unroll
                for (int k = 0; k < 5; ++k) Gs[tid * 5 + k] = gk[k];
            }
            const double Gxx = gk[0], Gxy = gk[1], Gyx = gk[2], Gyy = gk[3];
            if (graddiv) { Sxx = 2.0 * Gxx; Sxy = Gxy + Gyx; Syx = Sxy; Syy = 2.0 * Gyy; }
            else { Sxx = Gxx; Sxy = Gxy; Syx = Gyx; Syy = Gyy; }
            // cell term (:571): -int grad(psi):stress = +1/2 mean(nu) sum_j N_a,j S_ij
            const double hn = (1.0 / 6.0) * (nuv[0] + nuv[1] + nuv[2]);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                Rux[a] += hn * (Nx[a] * Sxx + Ny[a] * Sxy);
                Ruy[a] += hn * (Nx[a] * Syx + Ny[a] * Syy);
            }
            if (prm.graddepth) {
                // (:611-612): + int psi . (grad(H)/H . stress), non-polynomial: cell rule
                double gHx = 0, gHy = 0, hl[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    hl[a] = NONLIN ? b[a] + et[a] : b[a];
                    gHx += hl[a] * Nx[a];
                    gHy += hl[a] * Ny[a];
                }
                gHx *= -itA;
                gHy *= -itA;
                gdVx = gHx * Sxx + gHy * Syx;
                gdVy = gHx * Sxy + gHy * Syy;
#pragma unroll
                for (int a = 0; a < 3; ++a) gdnu[a] = nuv[a];
                // With Manning drag the cell-rule loop below already evaluates the depth and H^(-1/3) at every point:
                // the integrand k = w nu (dH/dhl) / H is accumulated there (gd_merge); otherwise here.
                if (!gd_merge) {
                    double Gd[3] = {0, 0, 0};
                    for (int qd = 0; qd < prm.nquad; ++qd) {
                        const double l0 = c_qlam[qd][0], l1 = c_qlam[qd][1], l2 = c_qlam[qd][2];
                        const double hq = l0 * hl[0] + l1 * hl[1] + l2 * hl[2];
                        double Hq = hq, fac = 1.0;
                        if (NONLIN && wd_on) {
                            // H = (hl + sqrt(hl^2 + alpha^2))/2, dH/dhl = (1 + hl/sqrt(hl^2 + alpha^2))/2
                            double a2q = a2;
                            if (var_al) {
                                const double aq = l0 * al[0] + l1 * al[1] + l2 * al[2];
                                a2q = aq * aq;
                                // NB grad(H) of the wetting-drying depth also has a d/d(alpha) part when alpha varies
                            }
                            const double x2 = fma(hq, hq, a2q);
                            const double r = tb_rsqrt(x2);
                            Hq = 0.5 * (hq + x2 * r);
                            fac = fma(0.5 * hq, r, 0.5);
                        }
                        const double k = c_qw[qd] * A * (l0 * nuv[0] + l1 * nuv[1] + l2 * nuv[2]) * fac * tb_rcp(Hq);
                        Gd[0] = fma(l0, k, Gd[0]); Gd[1] = fma(l1, k, Gd[1]); Gd[2] = fma(l2, k, Gd[2]);
                    }
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        Rux[a] = fma(Gd[a], gdVx, Rux[a]);
                        Ruy[a] = fma(Gd[a], gdVy, Ruy[a]);
                    }
                }
            }
        }
        if (use_quad) {
            // non-polynomial cell integrands by the degree-3 cell rule:
            // QuadraticDragTerm (:679-701), WindStressTerm (:643-649), wetting-drying HUDiv volume term
            double mu[3] = {0, 0, 0}, cdn[3] = {0, 0, 0}, twx[3] = {0, 0, 0}, twy[3] = {0, 0, 0};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (has_man) mu[a] = cf(prm.man, a);
                if (has_cd) cdn[a] = cf(prm.cd, a);
                if (has_nik) cdn[a] = cf(prm.nik, a);
                if (has_wind) {
                    twx[a] = cf(prm.wind, a, 0);
                    twy[a] = cf(prm.wind, a, 1);
                }
            }
            const double irho = has_wind ? tb_rcp(prm.rho0) : 0.0;
            double hl[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) hl[a] = NONLIN ? b[a] + et[a] : b[a];
            const int nq = SP::generic ? prm.nquad : 6;
            double HUx = 0.0, HUy = 0.0;
            double Gdm[3] = {0, 0, 0};
            const double gA = g * A;
            // Specialised kernels, symmetric rule (c_qsym): a P1 field f at point k is
            //   c sum_n f_n + (a - c) f_ia(k) + (b - c) f_ib(k) = fma(b - c, f_ib, P_ia),  P_n = fma(a - c, f_n, c sum_n f_n)
            // -- 10 operations per field for the six points instead of 18 -- and the common weight leaves the loop.
            const double qac = c_qsym[0], qbc = c_qsym[1], qc = c_qsym[2], qws = c_qsym[3];
            double Pu[3] = {0, 0, 0}, Pv[3] = {0, 0, 0}, Ph[3] = {0, 0, 0}, Pm[3] = {0, 0, 0}, Ptx[3] = {0, 0, 0},
                   Pty[3] = {0, 0, 0};
            if (!SP::generic) {
                const double cu = qc * sux, cv_ = qc * suy, ch = qc * (hl[0] + hl[1] + hl[2]);
                const double cm = qc * (mu[0] + mu[1] + mu[2]);
                const double ctx_ = qc * (twx[0] + twx[1] + twx[2]), cty_ = qc * (twy[0] + twy[1] + twy[2]);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    Pu[a] = fma(qac, ux[a], cu);
                    Pv[a] = fma(qac, uy[a], cv_);
                    Ph[a] = fma(qac, hl[a], ch);
                    Pm[a] = fma(qac, mu[a], cm);
                    Ptx[a] = fma(qac, twx[a], ctx_);
                    Pty[a] = fma(qac, twy[a], cty_);
                }
            }
TB_UNROLL(TB_QUAD_UNROLL)
            for (int qd = 0; qd < (SP::generic ? TB_MAX_QUAD : 6); ++qd) {
                if (SP::generic && qd >= nq) break;
                const double l0 = c_qlam[qd][0], l1 = c_qlam[qd][1], l2 = c_qlam[qd][2];
                const double wq = SP::generic ? c_qw[qd] : qws;
                const double w = wq * A;
                double uq, vq, Hq;
                if constexpr (SP::generic) {
                    uq = l0 * ux[0] + l1 * ux[1] + l2 * ux[2];
                    vq = l0 * uy[0] + l1 * uy[1] + l2 * uy[2];
                    Hq = l0 * hl[0] + l1 * hl[1] + l2 * hl[2];
                } else {
                    uq = fma(qbc, ux[tb_quad_ib(qd)], Pu[tb_quad_ia(qd)]);
                    vq = fma(qbc, uy[tb_quad_ib(qd)], Pv[tb_quad_ia(qd)]);
                    Hq = fma(qbc, hl[tb_quad_ib(qd)], Ph[tb_quad_ia(qd)]);
                }
                double gfac = 1.0;           // dH/dhl of the wetting-drying depth (grad-depth viscosity term)
                if (NONLIN) {
                    double a2q = a2;
                    if (var_al) {
                        const double aq = l0 * al[0] + l1 * al[1] + l2 * al[2];
                        a2q = aq * aq;
                    }
                    if (gd_merge && wd_on) {
                        const double x2 = fma(Hq, Hq, a2q);
                        const double rs = tb_rsqrt(x2);
                        gfac = fma(0.5 * Hq, rs, 0.5);
                        Hq = 0.5 * (Hq + x2 * rs);
                    } else {
                        Hq = wd_depth(Hq, wd_on, a2q);
                    }
                }
                double sx = 0, sy = 0;   // w * momentum source density at the point (the weight rides in the factors)
                if (has_man || has_cd || has_nik) {
                    const double s2 = fma(uq, uq, fma(vq, vq, prm.eps2));
                    const double umag = s2 > 0.0 ? tb_sqrt(s2) : 0.0;
                    double k;
                    if (has_man) {
                        double m;
                        if constexpr (SP::generic) m = l0 * mu[0] + l1 * mu[1] + l2 * mu[2];
                        else m = fma(qbc, mu[tb_quad_ib(qd)], Pm[tb_quad_ia(qd)]);
                        const double r = tb_rcbrt(Hq);             // H^(-1/3)
                        const double mr = m * (r * r);
                        k = (wq * -gA) * (mr * mr) * umag;         // -w g mu^2 / H^(1/3) * |u| / H = -w g (mu H^(-2/3))^2 |u|
                        if (gd_merge) {
                            // (:611-612) k = w nu (dH/dhl) / H with 1/H = (H^(-1/3))^3
                            const double kk = w * (l0 * gdnu[0] + l1 * gdnu[1] + l2 * gdnu[2]) * gfac * (r * r * r);
                            Gdm[0] = fma(l0, kk, Gdm[0]); Gdm[1] = fma(l1, kk, Gdm[1]); Gdm[2] = fma(l2, kk, Gdm[2]);
                        }
                    } else if (has_nik) {
                        // C_D = 2 kappa^2 / ln(11.036 H / k_s)^2 where H > k_s, else 0 (:697)
                        const double ks = l0 * cdn[0] + l1 * cdn[1] + l2 * cdn[2];
                        const double lg = log(11.036 * Hq / ks);
                        k = Hq > ks ? -w * (2.0 * prm.kappa * prm.kappa / (lg * lg) * umag / Hq) : 0.0;
                    } else {
                        k = -w * (l0 * cdn[0] + l1 * cdn[1] + l2 * cdn[2]) * umag * tb_rcp(Hq);
                    }
                    sx = k * uq;      // k carries the sign of the drag
                    sy = k * vq;
                }
                if (has_wind) {
                    const double k = (w * irho) * tb_rcp(Hq);
                    double tx, ty;
                    if constexpr (SP::generic) {
                        tx = l0 * twx[0] + l1 * twx[1] + l2 * twx[2];
                        ty = l0 * twy[0] + l1 * twy[1] + l2 * twy[2];
                    } else {
                        tx = fma(qbc, twx[tb_quad_ib(qd)], Ptx[tb_quad_ia(qd)]);
                        ty = fma(qbc, twy[tb_quad_ib(qd)], Pty[tb_quad_ia(qd)]);
                    }
                    sx = fma(k, tx, sx);
                    sy = fma(k, ty, sy);
                }
                Rux[0] += l0 * sx; Rux[1] += l1 * sx; Rux[2] += l2 * sx;
                Ruy[0] += l0 * sy; Ruy[1] += l1 * sy; Ruy[2] += l2 * sy;
                if (NONLIN && wd_on) {
                    // int H u over the cell by the rule (the test-function gradients are constant: applied after the loop)
                    const double wH = SP::generic ? c_qw[qd] * Hq : Hq;      // symmetric rule: the weight is applied after the loop
                    HUx = fma(wH, uq, HUx);
                    HUy = fma(wH, vq, HUy);
                }
            }
            if (gd_merge) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    Rux[a] = fma(Gdm[a], gdVx, Rux[a]);
                    Ruy[a] = fma(Gdm[a], gdVy, Ruy[a]);
                }
            }
            if (NONLIN && wd_on) {
                // grad(phi_a).(int H u)  with A grad(phi_a) = -N_a/2
                const double hs = SP::generic ? -0.5 : -0.5 * qws;
                HUx *= hs;
                HUy *= hs;
#pragma unroll
                for (int a = 0; a < 3; ++a) Re[a] = fma(Nx[a], HUx, fma(Ny[a], HUy, Re[a]));
            }
        }

        // ---------------- facet terms, 2-point Gauss per facet ----------------
        cp_async_wait_group0();
        __syncthreads();      // every thread's halo copies have landed
        const double g_half = 0.5 * g, g_quarter = 0.25 * g, lf_quarter = 0.25 * prm.lf_sigma;
        if (has_visc) {
            // gradients of the halo cells, one thread each, then visible to everybody
            const uint16_t *hcv = reinterpret_cast<const uint16_t *>(blk + prm.pl.off_hcv);
            for (int h = tid; h * 9 < nh9; h += TB_P) {
                double xh[3], yh[3], gh5[5];
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                    xh[bb] = cols[hcv[h * 3 + bb]];
                    yh[bb] = cols[NV + hcv[h * 3 + bb]];
                }
                tb_cell_gradient(S + (TB_P + h) * 9, xh, yh, gh5);
#pragma unroll
                for (int k = 0; k < 5; ++k) Gs[(TB_P + h) * 5 + k] = gh5[k];
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int p = (i + 1) % 3, q = (i + 2) % 3;
            const double nxs = Nx[i], nys = Ny[i];
            const double len2 = nxs * nxs + nys * nys;
            const double il = tb_rsqrt(len2);
            const double len = len2 * il;
            const int code = cn[i];
            if (code >= 0) {
                const int lf = code & 3;
                const double *nr = S + (code >> 2) * 9;
                const int np_ = (lf + 2) % 3, nq_ = (lf + 1) % 3;   // neighbour nodes matching p and q
                const double uNxp = nr[2 * np_], uNyp = nr[2 * np_ + 1], eNp = nr[6 + np_];
                // nodal differences along the facet: trace(xi) = v_p + xi (v_q - v_p)
                const double dKx = ux[q] - ux[p], dKy = uy[q] - uy[p], dKe = et[q] - et[p], dKb = b[q] - b[p];
                const double dNx = nr[2 * nq_] - uNxp, dNy = nr[2 * nq_ + 1] - uNyp, dNe = nr[6 + nq_] - eNp;
                // SIPG (:575-590): sigma_max*len, (S^K + S^N).N and the facet integral D of nu*(u_K - u_N)/2
                double sigl = 0, Tx = 0, Ty = 0, vDx = 0, vDy = 0, nup_ = 0, nuq_ = 0;
                if (has_visc) {
                    const double *gn = Gs + (code >> 2) * 5;        // the neighbour's grad(u) and 2A
                    const double *go = Gs + ct * 5;                  // this cell's
                    const double Hxx = gn[0] + go[0], Hxy = gn[1] + go[1], Hyx = gn[2] + go[2], Hyy = gn[3] + go[3];
                    const double twoAN = gn[4];
                    nup_ = nu_at(p);
                    nuq_ = nu_at(q);
                    double Qxx, Qxy, Qyx, Qyy;                      // S^K + S^N
                    if (graddiv) { Qxx = 2.0 * Hxx; Qxy = Hxy + Hyx; Qyx = Qxy; Qyy = 2.0 * Hyy; }
                    else { Qxx = Hxx; Qxy = Hxy; Qyx = Hyx; Qyy = Hyy; }
                    Tx = Qxx * nxs + Qxy * nys;
                    Ty = Qyx * nxs + Qyy * nys;
                    // sigma = sipg*cp*|e|/A (cp = 3 for P1 triangles), max over both sides
                    sigl = 6.0 * prm.sipg * len2 * tb_rcp(fmin(Gs[ct * 5 + 4], twoAN));
                }
TB_UNROLL(TB_GP_UNROLL)
                for (int gp = 0; gp < 2; ++gp) {
                    const double xi = gp ? TB_XI2 : TB_XI1;
                    const double hq_ = 0.5 * xi, hp_ = 0.5 - hq_;      // Gauss weight 1/2 folded into the test functions
                    const double uKx = fma(xi, dKx, ux[p]), uKy = fma(xi, dKy, uy[p]);
                    const double eK = fma(xi, dKe, et[p]);
                    const double uNx = fma(xi, dNx, uNxp), uNy = fma(xi, dNy, uNyp);
                    const double eN = fma(xi, dNe, eNp);
                    const double bg = fma(xi, dKb, b[p]);
                    // everything below is written in FMA-minimal form (no reassociation is left to the compiler)
                    const double esum = eK + eN, ediff = eK - eN;
                    double gh, hh;           // g*hbar and hbar/2, hbar = avg(total depth)
                    if (NONLIN && wd_on) {
                        const double hlK = bg + eK, hlN = bg + eN;
                        double a2g = a2;
                        if (var_al) {
                            const double ag = fma(xi, al[q] - al[p], al[p]);      // alpha is continuous (P1)
                            a2g = ag * ag;
                        }
                        const double sig = (hlK + hlN) + (tb_sqrt(fma(hlK, hlK, a2g)) + tb_sqrt(fma(hlN, hlN, a2g)));   // 4*hbar
                        gh = g_quarter * sig;
                        hh = 0.125 * sig;
                    } else {
                        const double hbar = NONLIN ? fma(0.5, esum, bg) : bg;
                        gh = g * hbar;
                        hh = 0.5 * hbar;
                    }
                    const double c = tb_sqrt(gh);
                    const double dux = uKx - uNx, duy = uKy - uNy;
                    const double usx = uKx + uNx, usy = uKy + uNy;
                    const double dun = fma(dux, nxs, duy * nys);
                    const double usN = fma(usx, nxs, usy * nys);
                    // PG (:363-366): g*(avg(eta) + sqrt(h/g)*jump(u,n)) n
                    const double t = fma(c * il, dun, g_half * esum);
                    // HUDiv (:424-427): h*(avg(u) + sqrt(g/h)*jump(eta,n)).n
                    const double fe = fma(c * len, ediff, hh * usN);
                    double fx, fy;
                    if (NONLIN && adv_on) {
                        // advection (:480-488): avg(u) (u_K.n) + gamma (u_K - u_N);  u_K.N = (us.N + du.N)/2
                        const double hu = 0.25 * (usN + dun);
                        if (lf_on) {
                            const double gam = lf_quarter * fabs(usN);
                            fx = fma(t, nxs, fma(hu, usx, gam * dux));
                            fy = fma(t, nys, fma(hu, usy, gam * duy));
                        } else {
                            fx = fma(t, nxs, hu * usx);
                            fy = fma(t, nys, hu * usy);
                        }
                    } else {
                        fx = t * nxs;
                        fy = t * nys;
                    }
                    if (has_visc) {
                        const double nug = fma(xi, nuq_ - nup_, nup_);          // nu is continuous (P1): avg(nu) = nu
                        double vx = sigl * dux, vy = sigl * duy;
                        if (graddiv) {
                            const double dn = sigl * dun * il * il;
                            vx += dn * nxs;
                            vy += dn * nys;
                        }
                        fx += nug * (vx - 0.5 * Tx);
                        fy += nug * (vy - 0.5 * Ty);
                        vDx += 0.5 * nug * dux;
                        vDy += 0.5 * nug * duy;
                    }
                    // tested against phi_p and phi_q, straight into the residual (no separate flux accumulators)
                    Rux[p] = fma(-hp_, fx, Rux[p]); Ruy[p] = fma(-hp_, fy, Ruy[p]); Re[p] = fma(-hp_, fe, Re[p]);
                    Rux[q] = fma(-hq_, fx, Rux[q]); Ruy[q] = fma(-hq_, fy, Ruy[q]); Re[q] = fma(-hq_, fe, Re[q]);
                }
                if (has_visc) {
                    // -inner(avg(grad(psi)), stress_jump) (:588): all three nodes, d_j phi_a = -N_a,j/2A
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double nn = Nx[a] * nxs + Ny[a] * nys;
                        double tx = vDx * nn, ty = vDy * nn;
                        if (graddiv) {
                            const double nd = Nx[a] * vDx + Ny[a] * vDy;
                            tx += nxs * nd;
                            ty += nys * nd;
                        }
                        Rux[a] -= 0.5 * itA * tx;
                        Ruy[a] -= 0.5 * itA * ty;
                    }
                }
            } else {
                const int gb = -(code + 1);
                const int slot = __ldg(prm.bc.bf_slot + gb);
                const int op = prm.bc.slots[slot].opcode;
                const bool closed = (op & (TB_BC_ELEV | TB_BC_UV | TB_BC_UN | TB_BC_FLUX)) == 0;
                double bDx = 0, bDy = 0;
                double Fpx = 0, Fpy = 0, Fpe = 0, Fqx = 0, Fqy = 0, Fqe = 0;   // boundary flux tested against phi_p, phi_q
#pragma unroll 1
                for (int gp = 0; gp < 2; ++gp) {
                    const double wq_ = gp ? TB_XI2 : TB_XI1, wp_ = 1.0 - wq_;
                    const double uKx = wp_ * ux[p] + wq_ * ux[q], uKy = wp_ * uy[p] + wq_ * uy[q];
                    const double eK = wp_ * et[p] + wq_ * et[q];
                    const double bg = wp_ * b[p] + wq_ * b[q];
                    double a2b = a2;
                    if (var_al) {
                        const double ab = wp_ * al[p] + wq_ * al[q];
                        a2b = ab * ab;
                    }
                    const double HK = NONLIN ? wd_depth(bg + eK, wd_on, a2b) : bg;
                    double fl[6];
                    if (closed) {
                        // land boundary (:376-381), mirror-velocity Lax-Friedrichs (:489-497)
                        const double uKN = uKx * nxs + uKy * nys;
                        const double c = tb_sqrt(g * HK);
                        double t = g * eK + c * uKN * il;
                        if (NONLIN && adv_on && lf_on) t += prm.lf_sigma * fabs(uKN) * uKN * il * il;
                        fl[0] = t * nxs;
                        fl[1] = t * nys;
                        fl[2] = 0.0;
                    } else {
                        open_boundary_flux<NONLIN>(&prm.bc, gb, slot, wp_, wq_, uKx, uKy, eK, bg, HK, nxs, nys, il, len, g,
                                                   wd_on, a2b, adv_on ? 1 : 0, fl);
                        if (has_visc && (op & (TB_BC_UV | TB_BC_UN | TB_BC_FLUX))) {
                            // Dirichlet terms of the viscosity (:592-609); 'elev' alone leaves uv_ext = uv: skipped
                            double ddx, ddy;
                            if (op & TB_BC_UN) {
                                const double sn = ((uKx * nxs + uKy * nys) * il - fl[5]) * il;
                                ddx = sn * nxs;
                                ddy = sn * nys;
                            } else {
                                ddx = uKx - fl[3];
                                ddy = uKy - fl[4];
                            }
                            const double nug = wp_ * nu_at(p) + wq_ * nu_at(q);
                            const double *go = Gs + ct * 5;
                            double Sxx, Sxy, Syx, Syy;
                            if (graddiv) { Sxx = 2.0 * go[0]; Sxy = go[1] + go[2]; Syx = Sxy; Syy = 2.0 * go[3]; }
                            else { Sxx = go[0]; Sxy = go[1]; Syx = go[2]; Syy = go[3]; }
                            const double sl = 6.0 * prm.sipg * len2 * itA;      // sigma*len, own cell only
                            double vx = sl * ddx, vy = sl * ddy;
                            if (graddiv) {
                                const double dn = sl * (ddx * nxs + ddy * nys) * il * il;
                                vx += dn * nxs;
                                vy += dn * nys;
                            }
                            fl[0] += nug * (vx - (Sxx * nxs + Sxy * nys));
                            fl[1] += nug * (vy - (Syx * nxs + Syy * nys));
                            bDx += 0.5 * nug * ddx;
                            bDy += 0.5 * nug * ddy;
                        }
                    }
                    Fpx += 0.5 * wp_ * fl[0]; Fpy += 0.5 * wp_ * fl[1]; Fpe += 0.5 * wp_ * fl[2];
                    Fqx += 0.5 * wq_ * fl[0]; Fqy += 0.5 * wq_ * fl[1]; Fqe += 0.5 * wq_ * fl[2];
                }
                if (has_visc && !closed) {
                    // -inner(grad(psi), stress_jump) ds (:606)
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double nn = Nx[a] * nxs + Ny[a] * nys;
                        double tx = bDx * nn, ty = bDy * nn;
                        if (graddiv) {
                            const double nd = Nx[a] * bDx + Ny[a] * bDy;
                            tx += nxs * nd;
                            ty += nys * nd;
                        }
                        Rux[a] -= itA * tx;
                        Ruy[a] -= itA * ty;
                    }
                }
                if (SP::generic && (op & TB_BC_DRAG)) {
                    double Fd[4] = {0.0, 0.0, 0.0, 0.0};
                    boundary_drag_facet(prm.bc.slots[slot].value, ux[p], uy[p], ux[q], uy[q], nxs, nys, il, len, Fd);
                    Fpx += Fd[0]; Fpy += Fd[1]; Fqx += Fd[2]; Fqy += Fd[3];
                }
                Rux[p] -= Fpx; Ruy[p] -= Fpy; Re[p] -= Fpe;
                Rux[q] -= Fqx; Ruy[q] -= Fqy; Re[q] -= Fqe;
            }
        }

        // ---------------- P1 mass inverse (equation.py:99-105) and Shu-Osher update ----------------
        // M_K^-1 = (3/A)(4 I - 1 1^T)
        const double mi = 6.0 * tb_rcp(twoA) * prm.bdt, mi4 = 4.0 * mi;
        const double sRx = mi * (Rux[0] + Rux[1] + Rux[2]), sRy = mi * (Ruy[0] + Ruy[1] + Ruy[2]),
                     sRe = mi * (Re[0] + Re[1] + Re[2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            res[2 * a] = fma(mi4, Rux[a], fma(prm.a1, ux[a], -sRx));
            res[2 * a + 1] = fma(mi4, Ruy[a], fma(prm.a1, uy[a], -sRy));
            res[6 + a] = fma(mi4, Re[a], fma(prm.a1, et[a], -sRe));
        }
        if (prm.u0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) res[k] = fma(prm.a0, O[tid * 9 + k], res[k]);
        }
        if (SP::generic && NONLIN && prm.wd_mass) {
            // explicit step on the reference's wetting-drying mass functional int (eta + f(b + eta)) phi
            // (shallowwater_eq.py:917-920) instead of the plain mass: res[6..8] holds the plain-mass update, the
            // cell-local Newton solve of tb_wd_mass.cuh turns it into the displaced-mass one
            double e0[3] = {0.0, 0.0, 0.0};
            if (prm.u0) {
#pragma unroll
                for (int a = 0; a < 3; ++a) e0[a] = O[tid * 9 + 6 + a];
            }
            tb_wd_displaced_update(res + 6, prm.u0 ? prm.a0 : 0.0, e0, prm.a1, et, b, var_al ? al : nullptr, a2,
                                   &c_qlam[0][0], c_qw, prm.nquad);
        }
        if (prm.partials && active) {
            // fused print_state / volume diagnostics of the state this launch produces (same closed forms as
            // swe_integrals_partial): int f g = A/12 (sum_a f_a g_a + (sum f)(sum g))
            const double A12 = twoA * (1.0 / 24.0);
            const double se = res[6] + res[7] + res[8];
            const double sx = res[0] + res[2] + res[4], sy = res[1] + res[3] + res[5];
            ia[0] = A12 * (res[6] * res[6] + res[7] * res[7] + res[8] * res[8] + se * se);
            ia[1] = A12 * (res[0] * res[0] + res[2] * res[2] + res[4] * res[4] + sx * sx + res[1] * res[1] +
                           res[3] * res[3] + res[5] * res[5] + sy * sy);
            ia[2] = twoA * (1.0 / 6.0) * se;
            ia[3] = twoA * (1.0 / 6.0) * (se + b[0] + b[1] + b[2]);
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) O[tid * 9 + k] = active ? res[k] : 0.0;
    double *red = reinterpret_cast<double *>(blk + prm.pl.stride);      // [TB_P / 32][4] scratch behind the static block
    if (prm.partials) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ia[k] += __shfl_down_sync(0xffffffffu, ia[k], o);
        }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) red[(tid >> 5) * 4 + k] = ia[k];
        }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(prm.u_out + cell0 * 9, O, TB_P * 9 * sizeof(double));
        bulk_commit_wait_read();
    } else if (prm.partials && tid >= 32 && tid < 36) {
        const int k = tid - 32;          // fixed order: deterministic
        double r = red[k];
#pragma unroll
        for (int w = 1; w < TB_P / 32; ++w) r += red[w * 4 + k];
        prm.partials[(long long)patch * 4 + k] = r;
    }
    // fused halo push: the records the peers need go from shared memory straight into their ghost blocks
    if (bpatch) tb_fused_push<9>(prm.halo, prm.push_dst, O, prm.n_bpatch, epoch, tid);
}

// shared-memory layout of the stage kernel: barrier | S (own + halo records) | O | static block | reduction scratch |
// [cell gradients (SIPG terms)] | halo ids
__host__ __device__ inline size_t tb_ids_offset(const TbPatchLayout &pl) {
    return 16 + (size_t)(TB_P + pl.NH) * 72 + (size_t)TB_P * 72 + (size_t)pl.stride + (TB_P / 32) * 4 * sizeof(double) +
           (pl.off_hcv >= 0 ? (size_t)(TB_P + pl.NH) * 5 * sizeof(double) : 0);
}
size_t tb_swe_smem_bytes(const TbPatchLayout &pl) {
    return tb_ids_offset(pl) + (((size_t)pl.NH * sizeof(int) + 15) & ~(size_t)15);
}

template <bool NL, int SPEC>
static cudaError_t stage_attr() {
    return cudaFuncSetAttribute(swe_stage_kernel<NL, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

cudaError_t tb_kernels_init() {
    cudaError_t e;
    if ((e = stage_attr<true, 0>()) != cudaSuccess) return e;
    if ((e = stage_attr<true, 1>()) != cudaSuccess) return e;
    if ((e = stage_attr<true, 2>()) != cudaSuccess) return e;
    if ((e = stage_attr<true, 3>()) != cudaSuccess) return e;
    if ((e = stage_attr<true, 5>()) != cudaSuccess) return e;
    if ((e = stage_attr<true, 6>()) != cudaSuccess) return e;
    if ((e = stage_attr<false, 0>()) != cudaSuccess) return e;
    if ((e = stage_attr<false, 4>()) != cudaSuccess) return e;
    return stage_attr<false, 1>();
}

// which specialisation serves this parameter set (0 = generic)
int tb_swe_stage_spec(const TbSweParams &p, bool nonlinear) {
    const bool dg_coef = p.cor.mode == 3 || p.man.mode == 3 || p.lin.mode == 3 || p.wind.mode == 3;   // P1DG coefficient fields
    const bool rare = p.cd.mode || p.pa.mode >= 2 || p.msrc.mode || p.vsrc.mode || !p.adv_on || p.nik.mode || p.wda.mode == 2 ||
                      dg_coef;
    if (rare || p.wd_mass) return 0;
    for (int j = 0; j < p.bc.n_slots; ++j)
        if (p.bc.slots[j].opcode & TB_BC_DRAG) return 0;     // BoundaryDragTerm lives in the generic kernel only
    // the specialised kernels with a cell rule are written for the symmetric 6-point rule (c_qsym)
    const bool quad6 = p.nquad == 6 && g_quad_sym;
    if (p.visc.mode) {
        if (nonlinear && p.lf_on && p.man.mode && p.cor.mode && !p.lin.mode && !p.wind.mode && quad6)
            return p.wd_on ? 6 : 5;
        return 0;
    }
    if (!nonlinear && p.cor.mode && p.wind.mode && p.lin.mode && !p.man.mode && !p.wd_on && quad6) return 4;
    if (p.lin.mode || p.wind.mode) return 0;
    if (!p.man.mode && !p.cor.mode && !p.wd_on && (!nonlinear || p.lf_on)) return 1;
    if (nonlinear && p.lf_on && p.man.mode && p.cor.mode && quad6) return p.wd_on ? 3 : 2;
    return 0;
}

cudaError_t tb_launch_swe_stage(const TbSweParams &p, bool nonlinear, int n_patches, size_t smem, cudaStream_t s) {
    if (n_patches <= 0) return cudaSuccess;
    const int spec = p.force_generic ? 0 : tb_swe_stage_spec(p, nonlinear);
    if (nonlinear) {
        switch (spec) {
            case 1: swe_stage_kernel<true, 1><<<n_patches, TB_P, smem, s>>>(p); break;
            case 2: swe_stage_kernel<true, 2><<<n_patches, TB_P, smem, s>>>(p); break;
            case 3: swe_stage_kernel<true, 3><<<n_patches, TB_P, smem, s>>>(p); break;
            case 5: swe_stage_kernel<true, 5><<<n_patches, TB_P, smem, s>>>(p); break;
            case 6: swe_stage_kernel<true, 6><<<n_patches, TB_P, smem, s>>>(p); break;
            default: swe_stage_kernel<true, 0><<<n_patches, TB_P, smem, s>>>(p); break;
        }
    } else {
        if (spec == 1) swe_stage_kernel<false, 1><<<n_patches, TB_P, smem, s>>>(p);
        else if (spec == 4) swe_stage_kernel<false, 4><<<n_patches, TB_P, smem, s>>>(p);
        else swe_stage_kernel<false, 0><<<n_patches, TB_P, smem, s>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t tb_set_quadrature(int n, const double *lam, const double *w) {
    cudaError_t e = cudaMemcpyToSymbol(c_qlam, lam, sizeof(double) * 3 * n);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_qw, w, sizeof(double) * n)) != cudaSuccess) return e;
    // symmetric structure the specialised kernels rely on: point k = (a at node ia(k), b at node ib(k), c at the third
    // node) for one triple (a, b, c), equal weights
    bool sym = (n == 6);
    double v[4] = {0, 0, 0, 0};
    if (sym) {
        const double a = lam[3 * 3 + 0], b = lam[3 * 3 + 1], c = lam[3 * 3 + 2];      // point 3 is (a, b, c)
        for (int k = 0; k < 6 && sym; ++k) {
            const int ia = tb_quad_ia(k), ib = tb_quad_ib(k), ic = 3 - ia - ib;
            sym = fabs(lam[3 * k + ia] - a) < 1e-14 && fabs(lam[3 * k + ib] - b) < 1e-14 && fabs(lam[3 * k + ic] - c) < 1e-14 &&
                  fabs(w[k] - w[0]) < 1e-15;
        }
        v[0] = a - c; v[1] = b - c; v[2] = c; v[3] = w[0];
    }
    g_quad_sym = sym;
    return cudaMemcpyToSymbol(c_qsym, v, sizeof(v));
}

// ------------------------------------------------------------------ layout conversion
__global__ void state_from_fields_kernel(const double *__restrict__ uv, const double *__restrict__ eta,
                                         const int32_t *__restrict__ node_map, double *__restrict__ state,
                                         long long n_cells) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (cell, node)
    if (i >= n_cells * 3) return;
    const long long c = i / 3;
    const int a = (int)(i - c * 3);
    const long long nd = node_map[i];
    state[c * 9 + 2 * a] = uv[2 * nd];
    state[c * 9 + 2 * a + 1] = uv[2 * nd + 1];
    state[c * 9 + 6 + a] = eta[nd];
}
__global__ void state_to_fields_kernel(const double *__restrict__ state, const int32_t *__restrict__ node_map,
                                       double *__restrict__ uv, double *__restrict__ eta, long long n_cells) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells * 3) return;
    const long long c = i / 3;
    const int a = (int)(i - c * 3);
    const long long nd = node_map[i];
    uv[2 * nd] = state[c * 9 + 2 * a];
    uv[2 * nd + 1] = state[c * 9 + 2 * a + 1];
    eta[nd] = state[c * 9 + 6 + a];
}
__global__ void tracer_from_field_kernel(const double *__restrict__ q, const int32_t *__restrict__ node_map,
                                         double *__restrict__ c, long long n3) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) c[i] = q[node_map[i]];
}
__global__ void tracer_to_field_kernel(const double *__restrict__ c, const int32_t *__restrict__ node_map,
                                       double *__restrict__ q, long long n3) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) q[node_map[i]] = c[i];
}
static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

cudaError_t tb_launch_state_from_fields(const double *uv, const double *eta, const int32_t *node_map, double *state,
                                        long long n_cells, cudaStream_t s) {
    if (n_cells) state_from_fields_kernel<<<nblk(n_cells * 3, 256), 256, 0, s>>>(uv, eta, node_map, state, n_cells);
    return cudaGetLastError();
}
cudaError_t tb_launch_state_to_fields(const double *state, const int32_t *node_map, double *uv, double *eta,
                                      long long n_cells, cudaStream_t s) {
    if (n_cells) state_to_fields_kernel<<<nblk(n_cells * 3, 256), 256, 0, s>>>(state, node_map, uv, eta, n_cells);
    return cudaGetLastError();
}
cudaError_t tb_launch_tracer_from_field(const double *q, const int32_t *node_map, double *c, long long n_cells,
                                        cudaStream_t s) {
    if (n_cells) tracer_from_field_kernel<<<nblk(n_cells * 3, 256), 256, 0, s>>>(q, node_map, c, n_cells * 3);
    return cudaGetLastError();
}
cudaError_t tb_launch_tracer_to_field(const double *c, const int32_t *node_map, double *q, long long n_cells,
                                      cudaStream_t s) {
    if (n_cells) tracer_to_field_kernel<<<nblk(n_cells * 3, 256), 256, 0, s>>>(c, node_map, q, n_cells * 3);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ halo pack / unpack
__global__ void gather_cells_kernel(const double *__restrict__ state, const int32_t *__restrict__ idx, long long n,
                                    int rec, double *__restrict__ buf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * rec) return;
    const long long h = i / rec;
    const int k = (int)(i - h * rec);
    buf[i] = state[(long long)idx[h] * rec + k];
}
__global__ void scatter_cells_kernel(const double *__restrict__ buf, const int32_t *__restrict__ idx, long long n,
                                     int rec, double *__restrict__ state) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * rec) return;
    const long long h = i / rec;
    const int k = (int)(i - h * rec);
    state[(long long)idx[h] * rec + k] = buf[i];
}
// peer push: record h goes straight to dst[h] (a pointer into a peer GPU's ghost block, NVLink P2P store)
__global__ void push_cells_kernel(const double *__restrict__ state, const int32_t *__restrict__ idx,
                                  const unsigned long long *__restrict__ dst, long long n, int rec) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * rec) return;
    const long long h = i / rec;
    const int k = (int)(i - h * rec);
    reinterpret_cast<double *>(dst[h])[k] = state[(long long)idx[h] * rec + k];
}
cudaError_t tb_launch_push_cells(const double *state, const int32_t *idx, const unsigned long long *dst, long long n,
                                 int rec, cudaStream_t s) {
    if (n) push_cells_kernel<<<nblk(n * rec, 256), 256, 0, s>>>(state, idx, dst, n, rec);
    return cudaGetLastError();
}
cudaError_t tb_launch_gather_cells(const double *state, const int32_t *idx, long long n, int rec, double *buf,
                                   cudaStream_t s) {
    if (n) gather_cells_kernel<<<nblk(n * rec, 256), 256, 0, s>>>(state, idx, n, rec, buf);
    return cudaGetLastError();
}
cudaError_t tb_launch_scatter_cells(const double *buf, const int32_t *idx, long long n, int rec, double *state,
                                    cudaStream_t s) {
    if (n) scatter_cells_kernel<<<nblk(n * rec, 256), 256, 0, s>>>(buf, idx, n, rec, state);
    return cudaGetLastError();
}

// Stream-ordered wait for the fused halo exchange: returns when every peer this rank receives from has published the
// ghost records of the last fused stage launch (needed in front of any OTHER kernel that reads those ghost records,
// e.g. the tracer stage reading the frozen SWE state).
__global__ void halo_fused_wait_kernel(const TbHaloFused *hf) {
    const unsigned long long epoch = *hf->epoch;
    const int q = threadIdx.x;
    if (q < hf->n_recv && !*reinterpret_cast<volatile int *>(hf->error)) {
        const unsigned long long *f = hf->flags + hf->recv_peer[q];
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < epoch) {
            if (clock64() - t0 > 6000000000ll) {
                *hf->error = 1;
                break;
            }
        }
    }
}
cudaError_t tb_launch_halo_fused_wait(const TbHaloFused *hf, cudaStream_t s) {
    halo_fused_wait_kernel<<<1, 32, 0, s>>>(hf);
    return cudaGetLastError();
}

// Rewrite the columns [col, col + ncomp) of every patch's static block from a P1 field given at the geometric
// vertices (vert[v*ncomp + comp]): stream-ordered refresh of a time-dependent coefficient (tb_sync_fields).
__global__ void update_columns_kernel(unsigned char *sblk, long long stride, int NV, long long n_patches,
                                      const int32_t *__restrict__ patch_vglob, const double *__restrict__ vert, int col,
                                      int ncomp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_patches * NV) return;
    const long long p = i / NV;
    const int k = (int)(i - p * NV);
    int gv = patch_vglob[i];
    if (gv < 0) gv = patch_vglob[p * NV];        // padding repeats the patch's first vertex (upload_layout)
    if (gv < 0) return;
    double *cols = reinterpret_cast<double *>(sblk + p * stride);
    for (int c = 0; c < ncomp; ++c) cols[(size_t)(col + c) * NV + k] = vert[(size_t)gv * ncomp + c];
}
cudaError_t tb_launch_update_columns(unsigned char *sblk, long long stride, int NV, long long n_patches,
                                     const int32_t *patch_vglob, const double *vert, int col, int ncomp, cudaStream_t s) {
    const long long n = n_patches * NV;
    if (n > 0) update_columns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sblk, stride, NV, n_patches, patch_vglob, vert, col, ncomp);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ Butcher-form stage combinations
// out = sum_j w[j] * x[j]   (ERKGeneric.update_solution / get_final_solution, rungekutta.py:816-852); streaming, HBM bound
struct TbLincomb {
    const double *x[6];
    double w[6];
    int n;
};
__global__ void lincomb_kernel(TbLincomb p, double *__restrict__ out, long long len2, int tail) {
    // two doubles per thread per iteration (16-byte accesses); an odd last element is handled by one thread
    if (tail && blockIdx.x == 0 && threadIdx.x == 0) {
        double acc = 0.0;
        for (int j = 0; j < p.n; ++j) acc = fma(p.w[j], p.x[j][2 * len2], acc);
        out[2 * len2] = acc;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len2; i += (long long)gridDim.x * blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (j < p.n) {
                const double2 v = reinterpret_cast<const double2 *>(p.x[j])[i];
                acc.x = fma(p.w[j], v.x, acc.x);
                acc.y = fma(p.w[j], v.y, acc.y);
            }
        }
        reinterpret_cast<double2 *>(out)[i] = acc;
    }
}
cudaError_t tb_launch_lincomb(int n, const double *const *x, const double *w, double *out, long long len, cudaStream_t s) {
    if (len <= 0) return cudaSuccess;
    TbLincomb p;
    p.n = n;
    for (int j = 0; j < 6; ++j) {
        p.x[j] = j < n ? x[j] : nullptr;
        p.w[j] = j < n ? w[j] : 0.0;
    }
    const long long len2 = len / 2;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((len2 + 255) / 256, 148 * 16));
    lincomb_kernel<<<grid, 256, 0, s>>>(p, out, len2, (int)(len & 1));
    return cudaGetLastError();
}

// ------------------------------------------------------------------ diagnostics
// Deterministic two-pass reductions (fixed grid, fixed order): TB_NRED CTAs write partials, one CTA sums them.
template <int NV_>
__device__ __forceinline__ void block_reduce_store(double (&a)[NV_], double *partial, const int *op) {
    __shared__ double sh[NV_][8];
#pragma unroll
    for (int k = 0; k < NV_; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            const double t = __shfl_down_sync(0xffffffffu, a[k], o);
            a[k] = op[k] == 0 ? a[k] + t : (op[k] == 1 ? fmin(a[k], t) : fmax(a[k], t));
        }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0)
#pragma unroll
        for (int k = 0; k < NV_; ++k) sh[k][w] = a[k];
    __syncthreads();
    if (threadIdx.x < NV_) {
        const int k = threadIdx.x;
        double r = sh[k][0];
        for (int j = 1; j < (int)(blockDim.x >> 5); ++j)
            r = op[k] == 0 ? r + sh[k][j] : (op[k] == 1 ? fmin(r, sh[k][j]) : fmax(r, sh[k][j]));
        partial[blockIdx.x * NV_ + k] = r;
    }
}
// out[0] = int eta^2, out[1] = int |u|^2, out[2] = int eta, out[3] = int (eta + bathymetry)  (comp_volume_2d)
__global__ void swe_integrals_partial(const double *__restrict__ state, const double *__restrict__ area,
                                      const double *__restrict__ bath3, long long n_owned, double *__restrict__ partial) {
    double a[4] = {0, 0, 0, 0};
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_owned;
         c += (long long)gridDim.x * blockDim.x) {
        const double *r = state + c * 9;
        const double A = area[c];
        // int f g = A/12 (sum_a f_a g_a + (sum f)(sum g))
        const double se = r[6] + r[7] + r[8];
        const double sx = r[0] + r[2] + r[4], sy = r[1] + r[3] + r[5];
        a[0] += A * (1.0 / 12.0) * (r[6] * r[6] + r[7] * r[7] + r[8] * r[8] + se * se);
        a[1] += A * (1.0 / 12.0) *
                (r[0] * r[0] + r[2] * r[2] + r[4] * r[4] + sx * sx + r[1] * r[1] + r[3] * r[3] + r[5] * r[5] + sy * sy);
        a[2] += A * (1.0 / 3.0) * se;
        a[3] += A * (1.0 / 3.0) * (se + bath3[c * 3] + bath3[c * 3 + 1] + bath3[c * 3 + 2]);
    }
    const int op[4] = {0, 0, 0, 0};
    block_reduce_store<4>(a, partial, op);
}
// out[0] = int c, out[1] = int H c (comp_tracer_mass_2d with H = total depth), out[2] = min c, out[3] = max c
__global__ void tracer_integrals_partial(const double *__restrict__ c, const double *__restrict__ swe,
                                         const double *__restrict__ area, const double *__restrict__ bath3,
                                         long long n_owned, int nonlin, int wd_on, double alpha2, int nquad,
                                         double *__restrict__ partial) {
    double a[4] = {0, 0, 1.0e300, -1.0e300};
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_owned;
         k += (long long)gridDim.x * blockDim.x) {
        const double q0 = c[k * 3], q1 = c[k * 3 + 1], q2 = c[k * 3 + 2];
        const double A = area[k];
        a[0] += A * (1.0 / 3.0) * (q0 + q1 + q2);
        double h[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) h[n] = bath3[k * 3 + n] + (nonlin ? swe[k * 9 + 6 + n] : 0.0);
        double m = 0;
        if (nonlin && wd_on) {
            // non-polynomial depth: the degree-3 cell rule (c_qlam / c_qw)
            for (int qd = 0; qd < nquad; ++qd) {
                const double l0 = c_qlam[qd][0], l1 = c_qlam[qd][1], l2 = c_qlam[qd][2];
                const double hq = l0 * h[0] + l1 * h[1] + l2 * h[2];
                m += c_qw[qd] * 0.5 * (hq + sqrt(hq * hq + alpha2)) * (l0 * q0 + l1 * q1 + l2 * q2);
            }
            m *= A;
        } else {
            m = A * (1.0 / 12.0) * (h[0] * q0 + h[1] * q1 + h[2] * q2 + (h[0] + h[1] + h[2]) * (q0 + q1 + q2));
        }
        a[1] += m;
        a[2] = fmin(a[2], fmin(q0, fmin(q1, q2)));
        a[3] = fmax(a[3], fmax(q0, fmax(q1, q2)));
    }
    const int op[4] = {0, 0, 1, 2};
    block_reduce_store<4>(a, partial, op);
}
__global__ void integrals_final(const double *__restrict__ partial, int nb, int op2, int op3, double *__restrict__ out) {
    // warp k reduces component k: lanes take the partials strided by 32, then a fixed shuffle tree (deterministic)
    const int k = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int op = k == 2 ? op2 : (k == 3 ? op3 : 0);
    double r = op == 0 ? 0.0 : (op == 1 ? 1.0e300 : -1.0e300);
    for (int j = l; j < nb; j += 32) {
        const double t = partial[j * 4 + k];
        r = op == 0 ? r + t : (op == 1 ? fmin(r, t) : fmax(r, t));
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_down_sync(0xffffffffu, r, o);
        r = op == 0 ? r + t : (op == 1 ? fmin(r, t) : fmax(r, t));
    }
    if (l == 0) out[k] = r;
}
// sum of per-patch partials [n][4] (written by the stage kernel's fused epilogue): one CTA, fixed order
__global__ void patch_partials_final(const double *__restrict__ partial, long long n, double *__restrict__ out) {
    __shared__ double sh[4][1024];
    double a[4] = {0, 0, 0, 0};
    for (long long j = threadIdx.x; j < n; j += blockDim.x) {
        const double4 v = reinterpret_cast<const double4 *>(partial)[j];
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
#pragma unroll
            for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) out[threadIdx.x] = sh[threadIdx.x][0];
}
cudaError_t tb_launch_patch_partials_final(const double *partial, long long n, double *out, cudaStream_t s) {
    patch_partials_final<<<1, 1024, 0, s>>>(partial, n, out);
    return cudaGetLastError();
}
cudaError_t tb_launch_swe_integrals(const double *state, const double *area, const double *bath3, long long n_owned,
                                    double *partial, double *out, cudaStream_t s) {
    swe_integrals_partial<<<TB_NRED, 256, 0, s>>>(state, area, bath3, n_owned, partial);
    integrals_final<<<1, 128, 0, s>>>(partial, TB_NRED, 0, 0, out);
    return cudaGetLastError();
}
cudaError_t tb_launch_tracer_integrals(const double *c, const double *swe, const double *area, const double *bath3,
                                       long long n_owned, int nonlin, int wd_on, double alpha2, int nquad,
                                       double *partial, double *out, cudaStream_t s) {
    tracer_integrals_partial<<<TB_NRED, 256, 0, s>>>(c, swe, area, bath3, n_owned, nonlin, wd_on, alpha2, nquad, partial);
    integrals_final<<<1, 128, 0, s>>>(partial, TB_NRED, 1, 2, out);
    return cudaGetLastError();
}
