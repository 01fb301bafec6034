// C-ABI glue: context, patch construction (native host runtime), option / coefficient /
// boundary-condition tables, kernel launches.  See include/thetis_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <nvtx3/nvToolsExt.h>
#include "tb_internal.h"

#define TB_BC_PRESENT 32

// NVTX range over one C-ABI call (header-only NVTX 3: a no-op unless a profiler is attached); the ranges name the
// phases of a step in an Nsight Systems timeline: tb_swe_stage[_fused], tb_tracer_stage, tb_limiter_apply,
// tb_push_cells, tb_sync_fields, tb_set_bc_array.
struct TbRange {
    explicit TbRange(const char *name) { nvtxRangePushA(name); }
    ~TbRange() { nvtxRangePop(); }
};

static thread_local std::string g_create_error;

struct FieldStore {
    int mode = 0;            // 0 none, 1 const, 2 vertex (P1), 3 cell nodes (P1DG)
    int ncomp = 1;
    double v[2] = {0, 0};
    std::vector<double> vert;   // [nv*ncomp]
    int col = -1;
    bool col_dirty = false;     // mode 2: `vert` changed, the columns of the static blocks are stale (tb_sync_fields)
    double *d_cell = nullptr;   // mode 3: DEVICE [n_owned_pad*3*ncomp]
    double *h_cell = nullptr;   // mode 3: pinned staging copy
    cudaEvent_t cell_event = nullptr;
};

struct tb_ctx {
    int device = 0;
    std::string err;
    // mesh (host copies)
    long long n_cells = 0, n_owned = 0, n_vertices = 0, n_bfacets = 0, n_tvert = 0;
    std::vector<double> coords;
    std::vector<int32_t> cells, nbr, bf_marker, topo;
    std::vector<int8_t> nbr_lf;
    long long n_patches = 0, n_owned_pad = 0;
    // patch tables (host)
    int NV = 0, NH = 0;
    std::vector<int32_t> patch_vglob;   // [n_patches*NV] global vertex of each patch-local vertex (-1 pad)
    std::vector<uint16_t> patch_cv;     // [n_patches*TB_P*3]
    std::vector<int32_t> patch_cn;      // [n_patches*TB_P*3]
    std::vector<int32_t> halo_ids;      // [n_patches*NH]
    std::vector<int32_t> halo_cnt;
    // device
    unsigned char *d_sblk = nullptr;
    size_t sblk_bytes = 0;
    int32_t *d_halo_ids = nullptr, *d_halo_cnt = nullptr;
    int32_t *d_bf_slot = nullptr, *d_bf_row = nullptr;
    std::vector<int32_t> bf_row;                 // compact row of each exterior facet (grouped by slot)
    std::vector<std::vector<int32_t>> slot_rows; // exterior facets of each slot, ascending
    std::vector<long long> slot_row0;            // first compact row of each slot
    // elev, uv, un, flux, value (swe), in TB_MAX_BANKS banks: bank i holds the boundary data of RK stage i when a whole
    // step is replayed from one CUDA graph (tb_set_bc_bank); everything else uses bank 0
    double *d_ext[TB_MAX_BANKS][5] = {};
    int bc_bank = 0;
    double *d_ext_tr[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int32_t *d_patch_vglob = nullptr;   // [n_patches*NV] device copy of patch_vglob (stream-ordered column updates)
    double *d_vert = nullptr, *h_vert = nullptr;     // staging of one vertex field (device / pinned), [n_vertices*2]
    cudaEvent_t vert_event = nullptr;
    double *d_area = nullptr;
    double *d_bath3 = nullptr;      // bathymetry at the 3 nodes of every owned cell (diagnostics)
    double *d_partial = nullptr;    // scratch of the two-pass reductions
    double *d_stage_partial = nullptr;  // [n_patches][4] written by the stage kernel's fused diagnostics epilogue
    int stage_integrals = 0;
    bool halo_geom = false;         // patch tables carry the halo cells' vertices (SIPG terms need the neighbour's gradient)
    std::vector<uint16_t> patch_hcv;    // [n_patches*NH*3]
    // ring of pinned staging buffers for boundary data uploads (no stream stall in steady state)
    static const int NSTAGE = 8;
    double *h_pinned[NSTAGE] = {nullptr};
    size_t h_pinned_bytes[NSTAGE] = {0};
    cudaEvent_t h_event[NSTAGE] = {nullptr};
    int h_next = 0;
    // layout
    TbPatchLayout pl{};
    bool layout_dirty = true;
    int ncol = 3;
    // options
    double g = 9.81, rho0 = 1000.0, lf_sigma = 1.0, norm_smoother = 0.0, wd_alpha = 0.5;
    int nonlinear = 1, lf_on = 1, wd_on = 0;
    int lf_tracer = 0;
    int force_generic = 0;
    double lf_tracer_sigma = 1.0, tracer_vel_factor = 1.0;
    double sipg = 1.0, sipg_tracer = 1.0, von_karman = 0.4;
    int graddiv = 0, graddepth = 1, tracer_conservative = 0, momentum_advection = 1;
    int wd_mass = 0;                // TB_OPT_WD_DISPLACED_MASS
    FieldStore fields[TB_F_COUNT];
    // bcs: eq 0 swe, 1 tracer
    std::vector<int> slot_marker;       // slot -> marker
    std::vector<int32_t> bf_slot;       // [n_bfacets]
    TbBcSlot bc[2][TB_MAX_SLOTS];
    // quadrature
    int nquad = 0;
    // patch range
    long long range_first = 0, range_count = -1;
    const int32_t *patch_list = nullptr;
    long long patch_list_n = 0;
    long long launches = 0;
    // fused halo exchange (tb_halo_fused_setup)
    bool fused_ready = false;
    int fused_n_bpatch = 0;
    TbHaloFused *d_fused = nullptr;
    int32_t *d_fused_order = nullptr, *d_push_ptr = nullptr, *d_push_cell = nullptr;
    unsigned long long *d_fused_epoch = nullptr;     // epoch (8 B), done counter (4 B), error flag (4 B)
    // limiter
    bool lim_ready = false;
    TbLimiterData lim{};
    unsigned char *d_lim_tab = nullptr;
    int32_t *d_lim_nhv = nullptr;
    double *d_lim_tmp = nullptr;
};

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                            \
            return TB_ERR_CUDA;                                                                        \
        }                                                                                              \
    } while (0)

static int fail(tb_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

// ------------------------------------------------------------------ patch construction
static int build_patches(tb_ctx *ctx) {
    const long long no = ctx->n_owned;
    const long long np = (no + TB_P - 1) / TB_P;
    ctx->n_patches = np;
    ctx->n_owned_pad = np * TB_P;
    // ghosts (cells >= n_owned) live after the padded owned block: cell id g -> g - n_owned + n_owned_pad
    auto dev_cell = [&](long long c) -> long long { return c < no ? c : c - no + ctx->n_owned_pad; };
    std::vector<int32_t> vstamp(ctx->n_vertices, -1), vlocal(ctx->n_vertices, 0);
    std::vector<int32_t> cstamp(ctx->n_cells, -1), clocal(ctx->n_cells, 0);
    std::vector<std::vector<int32_t>> pv(np), ph(np);
    int NV = 0, NH = 0;
    for (long long p = 0; p < np; ++p) {
        const long long c0 = p * TB_P, c1 = std::min(no, c0 + TB_P);
        auto &vl = pv[p];
        auto &hl = ph[p];
        for (long long c = c0; c < c1; ++c) {
            for (int a = 0; a < 3; ++a) {
                const int32_t gv = ctx->cells[c * 3 + a];
                if (gv < 0 || gv >= ctx->n_vertices) return fail(ctx, TB_ERR_ARG, "cell vertex id out of range");
                if (vstamp[gv] != p) {
                    vstamp[gv] = (int32_t)p;
                    vlocal[gv] = (int32_t)vl.size();
                    vl.push_back(gv);
                }
                const int32_t nb = ctx->nbr[c * 3 + a];
                if (nb >= 0) {
                    if (nb >= ctx->n_cells) return fail(ctx, TB_ERR_ARG, "neighbour id out of range");
                    if ((nb < c0 || nb >= c1) && cstamp[nb] != p) {
                        cstamp[nb] = (int32_t)p;
                        clocal[nb] = (int32_t)hl.size();
                        hl.push_back(nb);
                        if (ctx->halo_geom) {
                            for (int k = 0; k < 3; ++k) {
                                const int32_t hv = ctx->cells[(size_t)nb * 3 + k];
                                if (hv < 0 || hv >= ctx->n_vertices) return fail(ctx, TB_ERR_ARG, "cell vertex id out of range");
                                if (vstamp[hv] != p) {
                                    vstamp[hv] = (int32_t)p;
                                    vlocal[hv] = (int32_t)vl.size();
                                    vl.push_back(hv);
                                }
                            }
                        }
                    }
                } else if (nb == std::numeric_limits<int32_t>::min()) {
                    return fail(ctx, TB_ERR_ARG, "owned cell with unknown neighbour");
                } else if (-(long long)nb - 1 >= ctx->n_bfacets) {
                    return fail(ctx, TB_ERR_ARG, "exterior facet id out of range");
                }
            }
        }
        NV = std::max(NV, (int)vl.size());
        NH = std::max(NH, (int)hl.size());
    }
    NV = (NV + 1) & ~1;
    NH = (NH + 1) & ~1;
    if (NV > 65535) return fail(ctx, TB_ERR_UNSUPPORTED, "patch vertex table too large");
    ctx->NV = NV;
    ctx->NH = NH;
    ctx->patch_vglob.assign((size_t)np * NV, -1);
    ctx->patch_cv.assign((size_t)np * TB_P * 3, 0);
    ctx->patch_cn.assign((size_t)np * TB_P * 3, 0);
    ctx->halo_ids.assign((size_t)np * std::max(NH, 1), 0);
    ctx->halo_cnt.assign(np, 0);
    ctx->patch_hcv.assign(ctx->halo_geom ? (size_t)np * std::max(NH, 1) * 3 : 0, 0);
    // second pass: local ids (recompute stamps per patch)
    std::fill(vstamp.begin(), vstamp.end(), -1);
    std::fill(cstamp.begin(), cstamp.end(), -1);
    for (long long p = 0; p < np; ++p) {
        const long long c0 = p * TB_P, c1 = std::min(no, c0 + TB_P);
        const auto &vl = pv[p];
        const auto &hl = ph[p];
        for (size_t k = 0; k < vl.size(); ++k) {
            vstamp[vl[k]] = (int32_t)p;
            vlocal[vl[k]] = (int32_t)k;
            ctx->patch_vglob[(size_t)p * NV + k] = vl[k];
        }
        for (size_t k = 0; k < hl.size(); ++k) {
            cstamp[hl[k]] = (int32_t)p;
            clocal[hl[k]] = (int32_t)k;
            ctx->halo_ids[(size_t)p * NH + k] = (int32_t)dev_cell(hl[k]);
            if (ctx->halo_geom)
                for (int a = 0; a < 3; ++a)
                    ctx->patch_hcv[((size_t)p * NH + k) * 3 + a] = (uint16_t)vlocal[ctx->cells[(size_t)hl[k] * 3 + a]];
        }
        ctx->halo_cnt[p] = (int32_t)hl.size();
        for (long long c = c0; c < c1; ++c) {
            const size_t o = ((size_t)p * TB_P + (c - c0)) * 3;
            for (int a = 0; a < 3; ++a) {
                ctx->patch_cv[o + a] = (uint16_t)vlocal[ctx->cells[c * 3 + a]];
                const int32_t nb = ctx->nbr[c * 3 + a];
                if (nb >= 0) {
                    const int li = (nb >= c0 && nb < c1) ? (int)(nb - c0) : TB_P + clocal[nb];
                    ctx->patch_cn[o + a] = li * 4 + (int)ctx->nbr_lf[c * 3 + a];
                } else {
                    ctx->patch_cn[o + a] = nb;
                }
            }
        }
    }
    return TB_OK;
}

static int upload_halo_tables(tb_ctx *ctx) {
    if (ctx->d_halo_ids) cudaFree(ctx->d_halo_ids);
    if (ctx->d_halo_cnt) cudaFree(ctx->d_halo_cnt);
    ctx->d_halo_ids = ctx->d_halo_cnt = nullptr;
    CK(cudaMalloc(&ctx->d_halo_ids, sizeof(int32_t) * std::max<size_t>(ctx->halo_ids.size(), 4)));
    CK(cudaMemcpy(ctx->d_halo_ids, ctx->halo_ids.data(), sizeof(int32_t) * ctx->halo_ids.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ctx->d_halo_cnt, sizeof(int32_t) * ctx->n_patches));
    CK(cudaMemcpy(ctx->d_halo_cnt, ctx->halo_cnt.data(), sizeof(int32_t) * ctx->n_patches, cudaMemcpyHostToDevice));
    return TB_OK;
}

static int upload_layout(tb_ctx *ctx) {
    // SIPG terms need the geometry of the halo cells: rebuild the patch tables when that requirement changes
    const bool need_hgeom = ctx->fields[TB_F_VISCOSITY].mode != 0 || ctx->fields[TB_F_DIFFUSIVITY].mode != 0;
    // (sticky: several tracers with and without diffusivity share one context; switching back would rebuild the
    // tables on every change of owner)
    if (need_hgeom && !ctx->halo_geom) {
        CK(cudaDeviceSynchronize());
        ctx->halo_geom = true;
        int rc = build_patches(ctx);
        if (rc != TB_OK) return rc;
        rc = upload_halo_tables(ctx);
        if (rc != TB_OK) return rc;
    }
    // assign columns
    int ncol = 3;
    ctx->fields[TB_F_BATHYMETRY].col = 2;
    for (int f = 1; f < TB_F_COUNT; ++f) {
        FieldStore &fs = ctx->fields[f];
        fs.col = -1;
        if (fs.mode == 2) {
            fs.col = ncol;
            ncol += fs.ncomp;
        }
    }
    ctx->ncol = ncol;
    const int NV = ctx->NV;
    const long long np = ctx->n_patches;
    const size_t off_cv = (size_t)ncol * NV * sizeof(double);
    const size_t off_cn = off_cv + (size_t)TB_P * 3 * sizeof(uint16_t);
    const size_t off_hcv = off_cn + (size_t)TB_P * 3 * sizeof(int32_t);
    const size_t hcv_bytes = ctx->halo_geom ? (((size_t)ctx->NH * 3 * sizeof(uint16_t) + 15) & ~(size_t)15) : 0;
    const size_t stride = off_hcv + hcv_bytes;      // multiple of 16 (NV and NH are even): TMA bulk-copy granularity
    std::vector<unsigned char> host((size_t)np * stride, 0);
    const FieldStore &bath = ctx->fields[TB_F_BATHYMETRY];
    if (bath.mode == 0) return fail(ctx, TB_ERR_STATE, "bathymetry not set");
    for (long long p = 0; p < np; ++p) {
        unsigned char *blk = host.data() + (size_t)p * stride;
        double *cols = reinterpret_cast<double *>(blk);
        for (int k = 0; k < NV; ++k) {
            int gv = ctx->patch_vglob[(size_t)p * NV + k];
            if (gv < 0) gv = ctx->patch_vglob[(size_t)p * NV];   // pad with a valid vertex
            if (gv < 0) continue;
            cols[k] = ctx->coords[2 * (size_t)gv];
            cols[NV + k] = ctx->coords[2 * (size_t)gv + 1];
            cols[2 * NV + k] = bath.mode == 2 ? bath.vert[gv] : bath.v[0];
            for (int f = 1; f < TB_F_COUNT; ++f) {
                const FieldStore &fs = ctx->fields[f];
                if (fs.mode != 2) continue;
                for (int cpt = 0; cpt < fs.ncomp; ++cpt)
                    cols[(size_t)(fs.col + cpt) * NV + k] = fs.vert[(size_t)gv * fs.ncomp + cpt];
            }
        }
        memcpy(blk + off_cv, ctx->patch_cv.data() + (size_t)p * TB_P * 3, (size_t)TB_P * 3 * sizeof(uint16_t));
        memcpy(blk + off_cn, ctx->patch_cn.data() + (size_t)p * TB_P * 3, (size_t)TB_P * 3 * sizeof(int32_t));
        if (ctx->halo_geom)
            memcpy(blk + off_hcv, ctx->patch_hcv.data() + (size_t)p * ctx->NH * 3, (size_t)ctx->NH * 3 * sizeof(uint16_t));
    }
    // bathymetry at the cell nodes (volume / tracer-mass diagnostics)
    {
        std::vector<double> b3((size_t)ctx->n_owned * 3);
        for (long long c = 0; c < ctx->n_owned; ++c)
            for (int a = 0; a < 3; ++a)
                b3[(size_t)c * 3 + a] = bath.mode == 2 ? bath.vert[ctx->cells[(size_t)c * 3 + a]] : bath.v[0];
        if (!ctx->d_bath3) CK(cudaMalloc(&ctx->d_bath3, sizeof(double) * b3.size()));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ctx->d_bath3, b3.data(), sizeof(double) * b3.size(), cudaMemcpyHostToDevice));
    }
    if (host.size() != ctx->sblk_bytes) {
        if (ctx->d_sblk) cudaFree(ctx->d_sblk);
        ctx->d_sblk = nullptr;
        CK(cudaMalloc(&ctx->d_sblk, std::max<size_t>(host.size(), 16)));
        ctx->sblk_bytes = host.size();
    }
    // the blocks may be in use by kernels already queued: order the copy after them
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ctx->d_sblk, host.data(), host.size(), cudaMemcpyHostToDevice));
    ctx->pl.sblk = ctx->d_sblk;
    ctx->pl.stride = (long long)stride;
    ctx->pl.NV = NV;
    ctx->pl.NH = ctx->NH;
    ctx->pl.ncol = ncol;
    ctx->pl.off_cv = (int)off_cv;
    ctx->pl.off_cn = (int)off_cn;
    ctx->pl.off_hcv = ctx->halo_geom ? (int)off_hcv : -1;
    ctx->pl.halo_ids = ctx->d_halo_ids;
    ctx->pl.halo_cnt = ctx->d_halo_cnt;
    if (ctx->d_patch_vglob) cudaFree(ctx->d_patch_vglob);
    ctx->d_patch_vglob = nullptr;
    CK(cudaMalloc(&ctx->d_patch_vglob, sizeof(int32_t) * std::max<size_t>(ctx->patch_vglob.size(), 4)));
    CK(cudaMemcpy(ctx->d_patch_vglob, ctx->patch_vglob.data(), sizeof(int32_t) * ctx->patch_vglob.size(),
                  cudaMemcpyHostToDevice));
    for (int f = 0; f < TB_F_COUNT; ++f) ctx->fields[f].col_dirty = false;
    ctx->layout_dirty = false;
    return TB_OK;
}

// Stream-ordered refresh of coefficient data that changed since the last launch.  A P1 field whose VALUES changed
// (time-dependent wind stress / atmospheric pressure assigned in update_forcings) only rewrites its own columns of
// the per-patch static blocks: pinned staging copy -> async H2D -> one scatter kernel on `stream`; no device-wide
// synchronisation, so it is cheap every RK stage.  Structural changes (a field appearing / disappearing, bathymetry,
// the SIPG halo geometry) still rebuild the blocks (set-up time only).
static int sync_fields(tb_ctx *ctx, cudaStream_t st) {
    if (ctx->layout_dirty) {
        TbRange r("tb_sync_fields:rebuild_layout");
        return upload_layout(ctx);
    }
    for (int f = 1; f < TB_F_COUNT; ++f) {
        FieldStore &fs = ctx->fields[f];
        if (fs.mode != 2 || !fs.col_dirty) continue;
        TbRange r("tb_sync_fields:update_columns");
        const size_t n = (size_t)ctx->n_vertices * fs.ncomp;
        if (!ctx->d_vert) {
            CK(cudaMalloc(&ctx->d_vert, sizeof(double) * 2 * ctx->n_vertices));
            CK(cudaMallocHost(&ctx->h_vert, sizeof(double) * 2 * ctx->n_vertices));
            CK(cudaEventCreateWithFlags(&ctx->vert_event, cudaEventDisableTiming));
        } else {
            CK(cudaEventSynchronize(ctx->vert_event));       // the previous staging copy has been consumed
        }
        memcpy(ctx->h_vert, fs.vert.data(), sizeof(double) * n);
        CK(cudaMemcpyAsync(ctx->d_vert, ctx->h_vert, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        CK(tb_launch_update_columns(ctx->d_sblk, ctx->pl.stride, ctx->NV, ctx->n_patches, ctx->d_patch_vglob, ctx->d_vert,
                                    fs.col, fs.ncomp, st));
        CK(cudaEventRecord(ctx->vert_event, st));
        ctx->launches += 1;
        fs.col_dirty = false;
    }
    return TB_OK;
}

extern "C" int tb_sync_fields(tb_ctx *ctx, void *stream) {
    if (!ctx) return TB_ERR_ARG;
    return sync_fields(ctx, (cudaStream_t)stream);
}

static void default_quadrature(tb_ctx *ctx) {
    // Strang-Fix 6-point degree-3 rule (FIAT's classic default for degree 3)
    const double a = 0.659027622374092, b = 0.231933368553031, c = 0.109039009072877;
    const double xy[6][2] = {{a, b}, {a, c}, {b, a}, {b, c}, {c, a}, {c, b}};
    double lam[6][3], w[6];
    for (int i = 0; i < 6; ++i) {
        lam[i][0] = 1.0 - xy[i][0] - xy[i][1];
        lam[i][1] = xy[i][0];
        lam[i][2] = xy[i][1];
        w[i] = 1.0 / 6.0;
    }
    tb_set_quadrature(6, &lam[0][0], w);
    tb_set_quadrature_tracer(6, &lam[0][0], w);
    ctx->nquad = 6;
}

// ------------------------------------------------------------------ lifetime
extern "C" int tb_version(void) { return 100; }

extern "C" const char *tb_last_error(const tb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int tb_create(tb_ctx **out, const tb_mesh *m, int device) {
    if (!out || !m) return fail(nullptr, TB_ERR_ARG, "null argument");
    *out = nullptr;
    if (m->n_cells <= 0 || m->n_owned <= 0 || m->n_owned > m->n_cells || m->n_vertices <= 0 || !m->coords ||
        !m->cells || !m->nbr || !m->nbr_lf || (m->n_bfacets > 0 && !m->bf_marker))
        return fail(nullptr, TB_ERR_ARG, "invalid mesh description");
    if (m->n_cells > (1ll << 29)) return fail(nullptr, TB_ERR_UNSUPPORTED, "too many cells for 32-bit patch tables");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TB_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, TB_ERR_ARG, "bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, TB_ERR_CUDA, cudaGetErrorString(e));
    tb_ctx *ctx = new (std::nothrow) tb_ctx();
    if (!ctx) return fail(nullptr, TB_ERR_STATE, "out of host memory");
    ctx->device = device;
    ctx->n_cells = m->n_cells;
    ctx->n_owned = m->n_owned;
    ctx->n_vertices = m->n_vertices;
    ctx->n_bfacets = m->n_bfacets;
    ctx->coords.assign(m->coords, m->coords + 2 * m->n_vertices);
    ctx->cells.assign(m->cells, m->cells + 3 * m->n_cells);
    ctx->nbr.assign(m->nbr, m->nbr + 3 * m->n_cells);
    ctx->nbr_lf.assign(m->nbr_lf, m->nbr_lf + 3 * m->n_cells);
    if (m->n_bfacets) ctx->bf_marker.assign(m->bf_marker, m->bf_marker + m->n_bfacets);
    ctx->topo.resize(m->n_vertices);
    for (long long v = 0; v < m->n_vertices; ++v) ctx->topo[v] = m->topo ? m->topo[v] : (int32_t)v;
    ctx->n_tvert = 0;
    for (long long v = 0; v < m->n_vertices; ++v) ctx->n_tvert = std::max<long long>(ctx->n_tvert, ctx->topo[v] + 1);
    // geometry sanity: CCW cells
    std::vector<double> area(m->n_owned);
    for (long long c = 0; c < m->n_owned; ++c) {
        const double *p0 = &ctx->coords[2 * (size_t)ctx->cells[3 * c]];
        const double *p1 = &ctx->coords[2 * (size_t)ctx->cells[3 * c + 1]];
        const double *p2 = &ctx->coords[2 * (size_t)ctx->cells[3 * c + 2]];
        const double a2 = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p1[1] - p0[1]) * (p2[0] - p0[0]);
        if (!(a2 > 0)) {
            delete ctx;
            return fail(nullptr, TB_ERR_ARG, "cells must be counter-clockwise with positive area");
        }
        area[c] = 0.5 * a2;
    }
    int rc = build_patches(ctx);
    if (rc != TB_OK) {
        g_create_error = ctx->err;
        delete ctx;
        return rc;
    }
    // boundary marker slots
    ctx->bf_slot.resize(m->n_bfacets);
    for (long long k = 0; k < m->n_bfacets; ++k) {
        const int mk = ctx->bf_marker[k];
        int s = -1;
        for (size_t j = 0; j < ctx->slot_marker.size(); ++j)
            if (ctx->slot_marker[j] == mk) s = (int)j;
        if (s < 0) {
            if (ctx->slot_marker.size() >= TB_MAX_SLOTS) {
                delete ctx;
                return fail(nullptr, TB_ERR_UNSUPPORTED, "more than 16 distinct boundary markers");
            }
            s = (int)ctx->slot_marker.size();
            ctx->slot_marker.push_back(mk);
        }
        ctx->bf_slot[k] = s;
    }
    ctx->slot_rows.assign(TB_MAX_SLOTS, {});
    for (long long k = 0; k < m->n_bfacets; ++k) ctx->slot_rows[ctx->bf_slot[k]].push_back((int32_t)k);
    ctx->slot_row0.assign(TB_MAX_SLOTS + 1, 0);
    ctx->bf_row.assign(m->n_bfacets, 0);
    for (int sl = 0; sl < TB_MAX_SLOTS; ++sl) {
        ctx->slot_row0[sl + 1] = ctx->slot_row0[sl] + (long long)ctx->slot_rows[sl].size();
        for (size_t j = 0; j < ctx->slot_rows[sl].size(); ++j)
            ctx->bf_row[ctx->slot_rows[sl][j]] = (int32_t)(ctx->slot_row0[sl] + (long long)j);
    }
    memset(ctx->bc, 0, sizeof(ctx->bc));
    for (int eq = 0; eq < 2; ++eq)
        for (size_t j = 0; j < ctx->slot_marker.size(); ++j) {
            ctx->bc[eq][j].marker = ctx->slot_marker[j];
            ctx->bc[eq][j].bnd_len = 1.0;
        }
#define CKC(call)                                                              \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e__); \
            tb_destroy(ctx);                                                   \
            return TB_ERR_CUDA;                                                \
        }                                                                      \
    } while (0)
    CKC(tb_kernels_init());
    CKC(tb_tracer_kernels_init());
    if (upload_halo_tables(ctx) != TB_OK) {
        g_create_error = ctx->err;
        tb_destroy(ctx);
        return TB_ERR_CUDA;
    }
    CKC(cudaMalloc(&ctx->d_partial, sizeof(double) * 4 * TB_NRED));
    CKC(cudaMalloc(&ctx->d_bf_slot, sizeof(int32_t) * std::max<long long>(m->n_bfacets, 4)));
    if (m->n_bfacets)
        CKC(cudaMemcpy(ctx->d_bf_slot, ctx->bf_slot.data(), sizeof(int32_t) * m->n_bfacets, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&ctx->d_bf_row, sizeof(int32_t) * std::max<long long>(m->n_bfacets, 4)));
    if (m->n_bfacets)
        CKC(cudaMemcpy(ctx->d_bf_row, ctx->bf_row.data(), sizeof(int32_t) * m->n_bfacets, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&ctx->d_area, sizeof(double) * m->n_owned));
    CKC(cudaMemcpy(ctx->d_area, area.data(), sizeof(double) * m->n_owned, cudaMemcpyHostToDevice));
    default_quadrature(ctx);
    *out = ctx;
    return TB_OK;
}

extern "C" int tb_destroy(tb_ctx *ctx) {
    if (!ctx) return TB_OK;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_sblk);
    cudaFree(ctx->d_halo_ids);
    cudaFree(ctx->d_halo_cnt);
    cudaFree(ctx->d_bf_slot);
    cudaFree(ctx->d_bf_row);
    cudaFree(ctx->d_area);
    cudaFree(ctx->d_patch_vglob);
    cudaFree(ctx->d_vert);
    if (ctx->h_vert) cudaFreeHost(ctx->h_vert);
    if (ctx->vert_event) cudaEventDestroy(ctx->vert_event);
    for (int f = 0; f < TB_F_COUNT; ++f) {
        cudaFree(ctx->fields[f].d_cell);
        if (ctx->fields[f].h_cell) cudaFreeHost(ctx->fields[f].h_cell);
        if (ctx->fields[f].cell_event) cudaEventDestroy(ctx->fields[f].cell_event);
    }
    cudaFree(ctx->d_bath3);
    cudaFree(ctx->d_partial);
    cudaFree(ctx->d_stage_partial);
    for (int k = 0; k < 5; ++k) {
        for (int b = 0; b < TB_MAX_BANKS; ++b) cudaFree(ctx->d_ext[b][k]);
        cudaFree(ctx->d_ext_tr[k]);
    }
    cudaFree(ctx->d_fused);
    cudaFree(ctx->d_fused_order);
    cudaFree(ctx->d_push_ptr);
    cudaFree(ctx->d_push_cell);
    cudaFree(ctx->d_fused_epoch);
    cudaFree(ctx->d_lim_tab);
    cudaFree(ctx->d_lim_nhv);
    cudaFree(ctx->d_lim_tmp);
    for (int k = 0; k < tb_ctx::NSTAGE; ++k) {
        if (ctx->h_pinned[k]) cudaFreeHost(ctx->h_pinned[k]);
        if (ctx->h_event[k]) cudaEventDestroy(ctx->h_event[k]);
    }
    delete ctx;
    return TB_OK;
}

// ------------------------------------------------------------------ sizes
extern "C" int64_t tb_state_len(const tb_ctx *ctx) {
    return ctx ? (ctx->n_owned_pad + (ctx->n_cells - ctx->n_owned)) * 9 : 0;
}
extern "C" int64_t tb_tracer_len(const tb_ctx *ctx) {
    return ctx ? (ctx->n_owned_pad + (ctx->n_cells - ctx->n_owned)) * 3 : 0;
}
extern "C" int64_t tb_patch_size(const tb_ctx *) { return TB_P; }
extern "C" int64_t tb_n_patches(const tb_ctx *ctx) { return ctx ? ctx->n_patches : 0; }
extern "C" int64_t tb_launch_count(const tb_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------ options
extern "C" int tb_set_option(tb_ctx *ctx, int option, double value) {
    if (!ctx) return TB_ERR_ARG;
    switch (option) {
        case TB_OPT_G_GRAV: ctx->g = value; break;
        case TB_OPT_RHO0: ctx->rho0 = value; break;
        case TB_OPT_NONLINEAR: ctx->nonlinear = value != 0.0; break;
        case TB_OPT_LAX_FRIEDRICHS: ctx->lf_on = value != 0.0; break;
        case TB_OPT_LF_SCALING: ctx->lf_sigma = value; break;
        case TB_OPT_NORM_SMOOTHER: ctx->norm_smoother = value; break;
        case TB_OPT_WETTING_DRYING: ctx->wd_on = value != 0.0; break;
        case TB_OPT_WD_ALPHA: ctx->wd_alpha = value; break;
        case TB_OPT_LF_TRACER: ctx->lf_tracer = value != 0.0; break;
        case TB_OPT_LF_TRACER_SCALING: ctx->lf_tracer_sigma = value; break;
        case TB_OPT_TRACER_VEL_FACTOR: ctx->tracer_vel_factor = value; break;
        case TB_OPT_FORCE_GENERIC_KERNEL: ctx->force_generic = value != 0.0; break;
        case TB_OPT_SIPG_FACTOR: ctx->sipg = value; break;
        case TB_OPT_SIPG_FACTOR_TRACER: ctx->sipg_tracer = value; break;
        case TB_OPT_GRAD_DIV_VISCOSITY: ctx->graddiv = value != 0.0; break;
        case TB_OPT_GRAD_DEPTH_VISCOSITY: ctx->graddepth = value != 0.0; break;
        case TB_OPT_TRACER_CONSERVATIVE: ctx->tracer_conservative = value != 0.0; break;
        case TB_OPT_MOMENTUM_ADVECTION: ctx->momentum_advection = value != 0.0; break;
        case TB_OPT_VON_KARMAN: ctx->von_karman = value; break;
        case TB_OPT_WD_DISPLACED_MASS: ctx->wd_mass = value != 0.0; break;
        default: return fail(ctx, TB_ERR_ARG, "unknown option");
    }
    return TB_OK;
}

static int field_ncomp(int field) { return (field == TB_F_WIND_STRESS || field == TB_F_MOMENTUM_SOURCE) ? 2 : 1; }

extern "C" int tb_set_field_const(tb_ctx *ctx, int field, const double *value, int ncomp) {
    if (!ctx || !value || field < 0 || field >= TB_F_COUNT) return fail(ctx, TB_ERR_ARG, "bad field");
    if (ncomp != field_ncomp(field)) return fail(ctx, TB_ERR_ARG, "wrong number of components");
    FieldStore &fs = ctx->fields[field];
    if (fs.mode == 2 || field == TB_F_BATHYMETRY) ctx->layout_dirty = true;
    if (fs.mode == 0 && (field == TB_F_VISCOSITY || field == TB_F_DIFFUSIVITY)) ctx->layout_dirty = true;
    fs.mode = 1;
    fs.ncomp = ncomp;
    fs.v[0] = value[0];
    fs.v[1] = ncomp > 1 ? value[1] : 0.0;
    fs.vert.clear();
    return TB_OK;
}

extern "C" int tb_set_field_vertex(tb_ctx *ctx, int field, const double *values, int ncomp) {
    if (!ctx || !values || field < 0 || field >= TB_F_COUNT) return fail(ctx, TB_ERR_ARG, "bad field");
    if (ncomp != field_ncomp(field)) return fail(ctx, TB_ERR_ARG, "wrong number of components");
    FieldStore &fs = ctx->fields[field];
    // same field, same shape, new values: only its columns are refreshed (sync_fields); bathymetry also feeds the
    // diagnostics tables and a new field changes the block layout: full rebuild
    if (fs.mode == 2 && fs.ncomp == ncomp && field != TB_F_BATHYMETRY && !ctx->layout_dirty) fs.col_dirty = true;
    else ctx->layout_dirty = true;
    fs.mode = 2;
    fs.ncomp = ncomp;
    fs.vert.assign(values, values + (size_t)ctx->n_vertices * ncomp);
    return TB_OK;
}

extern "C" int tb_set_field_cell(tb_ctx *ctx, int field, const double *values, int ncomp, void *stream) {
    if (!ctx || !values || field < 0 || field >= TB_F_COUNT) return fail(ctx, TB_ERR_ARG, "bad field");
    if (ncomp != field_ncomp(field)) return fail(ctx, TB_ERR_ARG, "wrong number of components");
    if (field == TB_F_BATHYMETRY || field == TB_F_VISCOSITY || field == TB_F_DIFFUSIVITY || field == TB_F_WD_ALPHA)
        return fail(ctx, TB_ERR_UNSUPPORTED,
                    "this coefficient enters facet terms and must be continuous (P1): discontinuous data are outside "
                    "the accelerated path");
    FieldStore &fs = ctx->fields[field];
    const size_t n_owned3 = (size_t)ctx->n_owned * 3 * ncomp, n_pad3 = (size_t)ctx->n_owned_pad * 3 * ncomp;
    if (fs.mode == 2) ctx->layout_dirty = true;
    if (fs.d_cell && fs.ncomp != ncomp) {
        cudaFree(fs.d_cell);
        cudaFreeHost(fs.h_cell);
        fs.d_cell = fs.h_cell = nullptr;
    }
    if (!fs.d_cell) {
        CK(cudaMalloc(&fs.d_cell, sizeof(double) * n_pad3));
        CK(cudaMemset(fs.d_cell, 0, sizeof(double) * n_pad3));
        CK(cudaMallocHost(&fs.h_cell, sizeof(double) * n_owned3));
        if (!fs.cell_event) CK(cudaEventCreateWithFlags(&fs.cell_event, cudaEventDisableTiming));
    } else {
        CK(cudaEventSynchronize(fs.cell_event));
    }
    memcpy(fs.h_cell, values, sizeof(double) * n_owned3);
    CK(cudaMemcpyAsync(fs.d_cell, fs.h_cell, sizeof(double) * n_owned3, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CK(cudaEventRecord(fs.cell_event, (cudaStream_t)stream));
    fs.mode = 3;
    fs.ncomp = ncomp;
    fs.vert.clear();
    return TB_OK;
}

extern "C" int tb_clear_field(tb_ctx *ctx, int field) {
    if (!ctx || field < 0 || field >= TB_F_COUNT) return fail(ctx, TB_ERR_ARG, "bad field");
    FieldStore &fs = ctx->fields[field];
    if (fs.mode == 2) ctx->layout_dirty = true;
    if (fs.mode != 0 && (field == TB_F_VISCOSITY || field == TB_F_DIFFUSIVITY)) ctx->layout_dirty = true;
    fs.mode = 0;
    fs.vert.clear();
    return TB_OK;
}

static int find_slot(tb_ctx *ctx, int marker) {
    for (size_t j = 0; j < ctx->slot_marker.size(); ++j)
        if (ctx->slot_marker[j] == marker) return (int)j;
    return -1;
}

extern "C" int tb_set_bc(tb_ctx *ctx, int eq, int marker, int opcode, const double consts[8]) {
    if (!ctx || eq < 0 || eq > 1) return fail(ctx, TB_ERR_ARG, "bad equation id");
    const int s = find_slot(ctx, marker);
    if (s < 0) return TB_OK;   // marker not present on this (sub)mesh: nothing to do (reference loops over mesh markers)
    if (opcode & ~(TB_BC_ELEV | TB_BC_UV | TB_BC_UN | TB_BC_FLUX | TB_BC_VALUE | TB_BC_DIFF_FLUX | TB_BC_DRAG))
        return fail(ctx, TB_ERR_ARG, "invalid boundary tag");
    if ((opcode & TB_BC_DIFF_FLUX) && eq != 1) return fail(ctx, TB_ERR_ARG, "'diff_flux' is a tracer boundary tag");
    if ((opcode & TB_BC_DRAG) && eq != 0) return fail(ctx, TB_ERR_ARG, "'drag' is a shallow-water boundary tag");
    if ((opcode & TB_BC_DRAG) && !consts) return fail(ctx, TB_ERR_ARG, "'drag' needs its coefficient in consts[7]");
    TbBcSlot &b = ctx->bc[eq][s];
    b.opcode = opcode | TB_BC_PRESENT;
    b.arr_mask = 0;
    if (consts) {
        b.elev = consts[0]; b.uvx = consts[1]; b.uvy = consts[2];
        b.un = consts[3]; b.flux = consts[4]; b.value = consts[5];
        b.diff_flux = consts[6];
        // shallow-water slots have no 'value' datum (a tracer tag): the field carries the 'drag' coefficient there
        if (eq == 0) b.value = (opcode & TB_BC_DRAG) ? consts[7] : 0.0;
    }
    return TB_OK;
}

extern "C" int tb_set_bc_bank(tb_ctx *ctx, int bank) {
    if (!ctx || bank < 0 || bank >= TB_MAX_BANKS) return fail(ctx, TB_ERR_ARG, "bad boundary-data bank");
    ctx->bc_bank = bank;
    return TB_OK;
}

extern "C" int tb_clear_bc(tb_ctx *ctx, int eq, int marker) {
    if (!ctx || eq < 0 || eq > 1) return fail(ctx, TB_ERR_ARG, "bad equation id");
    const int s = find_slot(ctx, marker);
    if (s < 0) return TB_OK;
    TbBcSlot &b = ctx->bc[eq][s];
    b.opcode = 0;            // not even TB_BC_PRESENT: the marker has no entry in bnd_conditions (closed boundary)
    b.arr_mask = 0;
    b.elev = b.uvx = b.uvy = b.un = b.flux = b.value = b.diff_flux = 0.0;
    return TB_OK;
}

extern "C" int tb_set_boundary_length(tb_ctx *ctx, int marker, double length) {
    if (!ctx) return TB_ERR_ARG;
    const int s = find_slot(ctx, marker);
    if (s < 0) return TB_OK;
    ctx->bc[0][s].bnd_len = length;
    ctx->bc[1][s].bnd_len = length;
    return TB_OK;
}

extern "C" int tb_set_bc_array(tb_ctx *ctx, int eq, int marker, int tag, const double *values, int ncomp,
                               void *stream) {
    TbRange range("tb_set_bc_array");
    if (!ctx || eq < 0 || eq > 1 || !values) return fail(ctx, TB_ERR_ARG, "bad argument");
    const int s = find_slot(ctx, marker);
    if (s < 0) return TB_OK;
    int k;
    switch (tag) {
        case TB_BC_ELEV: k = 0; break;
        case TB_BC_UV: k = 1; break;
        case TB_BC_UN: k = 2; break;
        case TB_BC_FLUX: k = 3; break;
        case TB_BC_VALUE: k = 4; break;
        default: return fail(ctx, TB_ERR_ARG, "invalid boundary tag");
    }
    const int nc = (tag == TB_BC_UV) ? 2 : 1;
    if (ncomp != nc) return fail(ctx, TB_ERR_ARG, "wrong number of components");
    if (!(ctx->bc[eq][s].opcode & tag)) return fail(ctx, TB_ERR_STATE, "tag not declared with tb_set_bc");
    double **slot_arr = eq == 0 ? ctx->d_ext[ctx->bc_bank] : ctx->d_ext_tr;
    const size_t n = (size_t)ctx->n_bfacets * 2 * nc;
    if (!slot_arr[k]) {
        CK(cudaMalloc(&slot_arr[k], sizeof(double) * std::max<size_t>(n, 2)));
        if (eq == 0 && ctx->bc_bank > 0 && ctx->d_ext[0][k]) {
            // a new bank starts as a copy of bank 0: data of other markers / tags that do not change inside a step
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(slot_arr[k], ctx->d_ext[0][k], sizeof(double) * std::max<size_t>(n, 2), cudaMemcpyDeviceToDevice));
        } else {
            CK(cudaMemset(slot_arr[k], 0, sizeof(double) * std::max<size_t>(n, 2)));
        }
    }
    // The device arrays are compact: rows grouped by marker slot (bf_row), so one marker's data is one contiguous
    // block: gather its rows from the caller's full-size array into a pinned ring slot, one async copy.
    const size_t w = 2 * (size_t)nc;
    const std::vector<int32_t> &rows = ctx->slot_rows[s];
    if (rows.empty()) return TB_OK;
    const int slot = ctx->h_next;
    ctx->h_next = (ctx->h_next + 1) % tb_ctx::NSTAGE;
    if (!ctx->h_event[slot]) CK(cudaEventCreateWithFlags(&ctx->h_event[slot], cudaEventDisableTiming));
    else CK(cudaEventSynchronize(ctx->h_event[slot]));      // copy issued NSTAGE calls ago has finished
    const size_t bytes = sizeof(double) * w * rows.size();
    if (ctx->h_pinned_bytes[slot] < bytes) {
        if (ctx->h_pinned[slot]) cudaFreeHost(ctx->h_pinned[slot]);
        ctx->h_pinned[slot] = nullptr;
        CK(cudaMallocHost(&ctx->h_pinned[slot], bytes));
        ctx->h_pinned_bytes[slot] = bytes;
    }
    double *hp = ctx->h_pinned[slot];
    cudaStream_t st = (cudaStream_t)stream;
    if (w == 2) {
        for (size_t j = 0; j < rows.size(); ++j) {
            hp[2 * j] = values[2 * (size_t)rows[j]];
            hp[2 * j + 1] = values[2 * (size_t)rows[j] + 1];
        }
    } else {
        for (size_t j = 0; j < rows.size(); ++j) memcpy(hp + j * w, values + (size_t)rows[j] * w, sizeof(double) * w);
    }
    CK(cudaMemcpyAsync(slot_arr[k] + ctx->slot_row0[s] * w, hp, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->h_event[slot], st));
    ctx->bc[eq][s].arr_mask |= tag;
    return TB_OK;
}

// ------------------------------------------------------------------ hot path
static void fill_coef(const FieldStore &fs, TbCoef &c) {
    c.mode = fs.mode;
    c.col = fs.col;
    c.v0 = fs.v[0];
    c.v1 = fs.v[1];
    c.cell = fs.d_cell;
    c.nc = fs.ncomp;
}

static void fill_bc(tb_ctx *ctx, int eq, TbBcTable &t) {
    t.n_slots = (int)ctx->slot_marker.size();
    t.bf_slot = ctx->d_bf_slot;
    t.bf_row = ctx->d_bf_row;
    double **a = eq == 0 ? ctx->d_ext[ctx->bc_bank] : ctx->d_ext_tr;
    t.ext_elev = a[0];
    t.ext_uv = a[1];
    t.ext_un = a[2];
    t.ext_flux = a[3];
    t.ext_value = a[4];
    memcpy(t.slots, ctx->bc[eq], sizeof(TbBcSlot) * TB_MAX_SLOTS);
}

static void patch_range(tb_ctx *ctx, long long &first, long long &count) {
    first = 0;
    count = ctx->n_patches;
    if (ctx->range_count >= 0) {
        first = std::min(ctx->range_first, ctx->n_patches);
        count = std::min(ctx->range_count, ctx->n_patches - first);
    }
}

static int swe_stage_impl(tb_ctx *ctx, double a0, double a1, double b_dt, const double *u_in, const double *u0,
                          double *u_out, const unsigned long long *push_dst, void *stream) {
    TbRange range(push_dst ? "tb_swe_stage_fused" : "tb_swe_stage");
    if (!ctx || !u_in || !u_out) return fail(ctx, TB_ERR_ARG, "null state pointer");
    if (u_in == u_out) return fail(ctx, TB_ERR_ARG, "u_out must not alias u_in");
    if (a0 != 0.0 && !u0) return fail(ctx, TB_ERR_ARG, "u0 required when a0 != 0");
    if (ctx->fields[TB_F_MANNING].mode && ctx->fields[TB_F_QUAD_DRAG].mode)
        return fail(ctx, TB_ERR_ARG, "Cannot set both dimensionless and Manning drag parameter");
    {
        int rc = sync_fields(ctx, (cudaStream_t)stream);
        if (rc != TB_OK) return rc;
    }
    if ((ctx->fields[TB_F_MANNING].mode || ctx->fields[TB_F_QUAD_DRAG].mode) && ctx->fields[TB_F_NIKURADSE].mode)
        return fail(ctx, TB_ERR_ARG, "Cannot set both Nikuradse drag and Manning / dimensionless drag parameter");
    TbSweParams p;
    memset(&p, 0, sizeof(p));
    p.u_in = u_in;
    p.u0 = (a0 != 0.0) ? u0 : nullptr;
    p.u_out = u_out;
    p.pl = ctx->pl;
    p.n_owned = (int)ctx->n_owned;
    p.a0 = a0;
    p.a1 = a1;
    p.bdt = b_dt;
    p.g = ctx->g;
    p.rho0 = ctx->rho0;
    p.lf_sigma = ctx->lf_sigma;
    p.eps2 = ctx->norm_smoother * ctx->norm_smoother;
    p.wd_alpha2 = ctx->wd_alpha * ctx->wd_alpha;
    p.lf_on = ctx->lf_on;
    p.wd_on = ctx->wd_on && ctx->nonlinear;
    p.wd_mass = (ctx->wd_mass && p.wd_on) ? 1 : 0;
    if (p.wd_mass && fabs(a0 + a1 - 1.0) > 1e-9)
        return fail(ctx, TB_ERR_STATE,
                    "TB_OPT_WD_DISPLACED_MASS advances a mass functional: it needs a Shu-Osher stage (a0 + a1 = 1), not a "
                    "tendency evaluation");
    fill_coef(ctx->fields[TB_F_CORIOLIS], p.cor);
    fill_coef(ctx->fields[TB_F_MANNING], p.man);
    fill_coef(ctx->fields[TB_F_QUAD_DRAG], p.cd);
    fill_coef(ctx->fields[TB_F_LINEAR_DRAG], p.lin);
    fill_coef(ctx->fields[TB_F_WIND_STRESS], p.wind);
    fill_coef(ctx->fields[TB_F_ATM_PRESSURE], p.pa);
    fill_coef(ctx->fields[TB_F_MOMENTUM_SOURCE], p.msrc);
    fill_coef(ctx->fields[TB_F_VOLUME_SOURCE], p.vsrc);
    fill_coef(ctx->fields[TB_F_VISCOSITY], p.visc);
    fill_coef(ctx->fields[TB_F_NIKURADSE], p.nik);
    fill_coef(ctx->fields[TB_F_WD_ALPHA], p.wda);
    if (p.wda.mode == 1) {       // a Constant given as a field: same as the scalar option
        p.wd_alpha2 = p.wda.v0 * p.wda.v0;
        p.wda.mode = 0;
    }
    p.kappa = ctx->von_karman;
    p.sipg = ctx->sipg;
    p.graddiv = ctx->graddiv;
    p.graddepth = ctx->graddepth;
    p.use_quad = (p.man.mode || p.cd.mode || p.nik.mode || p.wind.mode || p.wd_on) ? 1 : 0;
    p.nquad = ctx->nquad;
    p.force_generic = ctx->force_generic;
    p.adv_on = ctx->momentum_advection;
    p.partials = nullptr;
    if (ctx->stage_integrals) {
        if (!ctx->d_stage_partial) {
            CK(cudaMalloc(&ctx->d_stage_partial, sizeof(double) * 4 * ctx->n_patches));
            CK(cudaMemset(ctx->d_stage_partial, 0, sizeof(double) * 4 * ctx->n_patches));
        }
        p.partials = ctx->d_stage_partial;
    }
    fill_bc(ctx, 0, p.bc);
    long long first, count;
    patch_range(ctx, first, count);
    p.patch_first = (int)first;
    if (ctx->patch_list) {
        p.patch_list = ctx->patch_list;
        p.patch_first = 0;
        count = ctx->patch_list_n;
    }
    if (push_dst) {
        // one launch over all patches, partition-boundary patches first; they push their records to the peers
        if (!ctx->fused_ready) return fail(ctx, TB_ERR_STATE, "tb_halo_fused_setup has not been called");
        p.patch_list = ctx->d_fused_order;
        p.patch_first = 0;
        count = ctx->n_patches;
        p.halo = ctx->d_fused;
        p.push_dst = push_dst;
        p.n_bpatch = ctx->fused_n_bpatch;
    }
    const size_t smem = tb_swe_smem_bytes(ctx->pl);
    if (smem > 200 * 1024) return fail(ctx, TB_ERR_UNSUPPORTED, "patch halo too large for shared memory");
    CK(tb_launch_swe_stage(p, ctx->nonlinear != 0, (int)count, smem, (cudaStream_t)stream));
    ctx->launches += count > 0 ? 1 : 0;
    return TB_OK;
}

extern "C" int tb_swe_stage(tb_ctx *ctx, double a0, double a1, double b_dt, const double *u_in, const double *u0,
                            double *u_out, void *stream) {
    return swe_stage_impl(ctx, a0, a1, b_dt, u_in, u0, u_out, nullptr, stream);
}

extern "C" int tb_swe_stage_fused(tb_ctx *ctx, double a0, double a1, double b_dt, const double *u_in, const double *u0,
                                  double *u_out, const uint64_t *push_dst, void *stream) {
    if (!push_dst) return fail(ctx, TB_ERR_ARG, "null push table");
    return swe_stage_impl(ctx, a0, a1, b_dt, u_in, u0, u_out, reinterpret_cast<const unsigned long long *>(push_dst),
                          stream);
}

extern "C" int tb_halo_fused_setup(tb_ctx *ctx, const tb_halo_fused *h) {
    if (!ctx || !h) return fail(ctx, TB_ERR_ARG, "null argument");
    if (h->n_bpatch < 0 || h->n_bpatch > ctx->n_patches || !h->patch_order || !h->flags ||
        (h->n_bpatch > 0 && (!h->push_ptr || !h->push_cell)))
        return fail(ctx, TB_ERR_ARG, "invalid fused halo description");
    if (h->n_recv < 0 || h->n_recv > TB_MAX_PEERS || h->n_send < 0 || h->n_send > TB_MAX_PEERS)
        return fail(ctx, TB_ERR_UNSUPPORTED, "more than 16 neighbouring ranks");
    std::vector<char> seen(ctx->n_patches, 0);
    for (long long k = 0; k < ctx->n_patches; ++k) {
        const int32_t q = h->patch_order[k];
        if (q < 0 || q >= ctx->n_patches || seen[q]) return fail(ctx, TB_ERR_ARG, "patch_order is not a permutation");
        seen[q] = 1;
    }
    const long long ne = h->n_bpatch > 0 ? h->push_ptr[h->n_bpatch] : 0;
    for (long long e = 0; e < ne; ++e)
        if (h->push_cell[e] < 0 || h->push_cell[e] >= TB_P) return fail(ctx, TB_ERR_ARG, "push cell outside its patch");
    CK(cudaDeviceSynchronize());
    cudaFree(ctx->d_fused); cudaFree(ctx->d_fused_order); cudaFree(ctx->d_push_ptr); cudaFree(ctx->d_push_cell);
    cudaFree(ctx->d_fused_epoch);
    ctx->d_fused = nullptr; ctx->d_fused_order = ctx->d_push_ptr = ctx->d_push_cell = nullptr; ctx->d_fused_epoch = nullptr;
    ctx->fused_ready = false;
    CK(cudaMalloc(&ctx->d_fused_order, sizeof(int32_t) * ctx->n_patches));
    CK(cudaMemcpy(ctx->d_fused_order, h->patch_order, sizeof(int32_t) * ctx->n_patches, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ctx->d_push_ptr, sizeof(int32_t) * (h->n_bpatch + 1)));
    CK(cudaMalloc(&ctx->d_push_cell, sizeof(int32_t) * std::max<long long>(ne, 1)));
    if (h->n_bpatch > 0) {
        CK(cudaMemcpy(ctx->d_push_ptr, h->push_ptr, sizeof(int32_t) * (h->n_bpatch + 1), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->d_push_cell, h->push_cell, sizeof(int32_t) * ne, cudaMemcpyHostToDevice));
    } else {
        CK(cudaMemset(ctx->d_push_ptr, 0, sizeof(int32_t)));
    }
    CK(cudaMalloc(&ctx->d_fused_epoch, 16));
    CK(cudaMemset(ctx->d_fused_epoch, 0, 16));
    TbHaloFused hf;
    memset(&hf, 0, sizeof(hf));
    hf.epoch = ctx->d_fused_epoch;
    hf.done_count = reinterpret_cast<unsigned int *>(ctx->d_fused_epoch + 1);
    hf.error = reinterpret_cast<int *>(ctx->d_fused_epoch + 1) + 1;
    hf.flags = reinterpret_cast<const unsigned long long *>(h->flags);
    hf.push_ptr = ctx->d_push_ptr;
    hf.push_cell = ctx->d_push_cell;
    hf.n_recv = h->n_recv;
    hf.n_send = h->n_send;
    for (int q = 0; q < h->n_recv; ++q) hf.recv_peer[q] = h->recv_peer[q];
    for (int q = 0; q < h->n_send; ++q) hf.remote_flag[q] = reinterpret_cast<unsigned long long *>(h->remote_flag[q]);
    CK(cudaMalloc(&ctx->d_fused, sizeof(TbHaloFused)));
    CK(cudaMemcpy(ctx->d_fused, &hf, sizeof(hf), cudaMemcpyHostToDevice));
    ctx->fused_n_bpatch = (int)h->n_bpatch;
    ctx->fused_ready = true;
    return TB_OK;
}

extern "C" int tb_halo_fused_wait(tb_ctx *ctx, void *stream) {
    if (!ctx || !ctx->fused_ready) return fail(ctx, TB_ERR_STATE, "tb_halo_fused_setup has not been called");
    CK(tb_launch_halo_fused_wait(ctx->d_fused, (cudaStream_t)stream));
    ctx->launches += 1;
    return TB_OK;
}

extern "C" int tb_halo_fused_status(tb_ctx *ctx, int64_t *epoch, int32_t *error) {
    if (!ctx || !ctx->fused_ready) return fail(ctx, TB_ERR_STATE, "tb_halo_fused_setup has not been called");
    unsigned long long h[2];
    CK(cudaMemcpy(h, ctx->d_fused_epoch, 16, cudaMemcpyDeviceToHost));      // synchronises
    if (epoch) *epoch = (int64_t)h[0];
    if (error) *error = (int32_t)(h[1] >> 32);
    return TB_OK;
}

extern "C" int tb_swe_tendency(tb_ctx *ctx, const double *u, double *k_out, void *stream) {
    return tb_swe_stage(ctx, 0.0, 0.0, 1.0, u, nullptr, k_out, stream);
}

static int tracer_stage_impl(tb_ctx *ctx, double a0, double a1, double b_dt, const double *c_in, const double *c0,
                             double *c_out, const double *swe_state, const unsigned long long *push_dst, void *stream) {
    TbRange range(push_dst ? "tb_tracer_stage_fused" : "tb_tracer_stage");
    if (!ctx || !c_in || !c_out || !swe_state) return fail(ctx, TB_ERR_ARG, "null state pointer");
    if (c_in == c_out) return fail(ctx, TB_ERR_ARG, "c_out must not alias c_in");
    if (a0 != 0.0 && !c0) return fail(ctx, TB_ERR_ARG, "c0 required when a0 != 0");
    {
        int rc = sync_fields(ctx, (cudaStream_t)stream);
        if (rc != TB_OK) return rc;
    }
    TbTracerParams p;
    memset(&p, 0, sizeof(p));
    p.c_in = c_in;
    p.c0 = (a0 != 0.0) ? c0 : nullptr;
    p.c_out = c_out;
    p.swe = swe_state;
    p.pl = ctx->pl;
    p.n_owned = (int)ctx->n_owned;
    p.a0 = a0;
    p.a1 = a1;
    p.bdt = b_dt;
    p.corr = ctx->tracer_vel_factor;
    p.lf_sigma = ctx->lf_tracer_sigma;
    p.lf_on = ctx->lf_tracer;
    p.nonlin = ctx->nonlinear;
    p.wd_on = ctx->wd_on && ctx->nonlinear;
    p.wd_alpha2 = ctx->wd_alpha * ctx->wd_alpha;
    fill_coef(ctx->fields[TB_F_TRACER_SOURCE], p.src);
    fill_coef(ctx->fields[TB_F_DIFFUSIVITY], p.diff);
    p.sipg = ctx->sipg_tracer;
    p.conservative = ctx->tracer_conservative;
    p.nquad = ctx->nquad;
    p.force_generic = ctx->force_generic;
    fill_bc(ctx, 1, p.bc);
    long long first, count;
    patch_range(ctx, first, count);
    p.patch_first = (int)first;
    if (push_dst) {
        // one launch over all patches, partition-boundary patches first; they push their values to the peers
        if (!ctx->fused_ready) return fail(ctx, TB_ERR_STATE, "tb_halo_fused_setup has not been called");
        p.patch_list = ctx->d_fused_order;
        p.patch_first = 0;
        count = ctx->n_patches;
        p.halo = ctx->d_fused;
        p.push_dst = push_dst;
        p.n_bpatch = ctx->fused_n_bpatch;
    }
    const size_t smem = tb_tracer_smem_bytes(ctx->pl);
    if (smem > 200 * 1024) return fail(ctx, TB_ERR_UNSUPPORTED, "patch halo too large for shared memory");
    CK(tb_launch_tracer_stage(p, (int)count, smem, (cudaStream_t)stream));
    ctx->launches += count > 0 ? 1 : 0;
    return TB_OK;
}

extern "C" int tb_tracer_stage(tb_ctx *ctx, double a0, double a1, double b_dt, const double *c_in, const double *c0,
                               double *c_out, const double *swe_state, void *stream) {
    return tracer_stage_impl(ctx, a0, a1, b_dt, c_in, c0, c_out, swe_state, nullptr, stream);
}

extern "C" int tb_tracer_stage_fused(tb_ctx *ctx, double a0, double a1, double b_dt, const double *c_in, const double *c0,
                                     double *c_out, const double *swe_state, const uint64_t *push_dst, void *stream) {
    if (!push_dst) return fail(ctx, TB_ERR_ARG, "null push table");
    return tracer_stage_impl(ctx, a0, a1, b_dt, c_in, c0, c_out, swe_state,
                             reinterpret_cast<const unsigned long long *>(push_dst), stream);
}

static int limiter_setup(tb_ctx *ctx) {
    // per-patch tables of the patch-staged limiter kernel (tb_tracer.cu): vertex halo, patch-local topological
    // vertex ids of own and halo cells, exterior-facet masks
    const long long nc = ctx->n_cells, no = ctx->n_owned, nt = ctx->n_tvert, np = ctx->n_patches;
    auto dev_cell = [&](long long c) -> long long { return c < no ? c : c - no + ctx->n_owned_pad; };
    auto tvert = [&](long long c, int a) -> int { return ctx->topo[ctx->cells[3 * c + a]]; };
    auto bmask = [&](long long c) -> unsigned char {
        unsigned char m = 0;
        for (int f = 0; f < 3; ++f) {
            const int32_t nb = ctx->nbr[3 * c + f];
            if (nb < 0 && nb != std::numeric_limits<int32_t>::min()) m |= (unsigned char)(1 << f);
        }
        return m;
    };
    // cells around each topological vertex (all local cells, ghosts included)
    std::vector<long long> ptr(nt + 1, 0);
    for (long long c = 0; c < nc; ++c)
        for (int a = 0; a < 3; ++a) ptr[tvert(c, a) + 1]++;
    for (long long v = 0; v < nt; ++v) ptr[v + 1] += ptr[v];
    std::vector<int> idx(ptr[nt]);
    {
        std::vector<long long> fill(ptr.begin(), ptr.end() - 1);
        for (long long c = 0; c < nc; ++c)
            for (int a = 0; a < 3; ++a) idx[fill[tvert(c, a)]++] = (int)c;
    }
    std::vector<std::vector<int>> pverts(np), phalo(np);
    std::vector<int> vstamp(nt, -1), vloc(nt, 0), cstamp(nc, -1);
    int NVT = 0, NHV = 0;
    for (long long p = 0; p < np; ++p) {
        const long long c0 = p * TB_P, c1 = std::min(no, c0 + TB_P);
        for (long long c = c0; c < c1; ++c)
            for (int a = 0; a < 3; ++a) {
                const int v = tvert(c, a);
                if (vstamp[v] != p) {
                    vstamp[v] = (int)p;
                    pverts[p].push_back(v);
                    for (long long k = ptr[v]; k < ptr[v + 1]; ++k) {
                        const int h = idx[k];
                        if ((h < c0 || h >= c1) && cstamp[h] != p) {
                            cstamp[h] = (int)p;
                            phalo[p].push_back(h);
                        }
                    }
                }
            }
        NVT = std::max(NVT, (int)pverts[p].size());
        NHV = std::max(NHV, (int)phalo[p].size());
    }
    if (NHV + TB_P >= 8192) return fail(ctx, TB_ERR_UNSUPPORTED, "patch vertex halo too large");
    NVT = (NVT + 1) & ~1;
    NHV = std::max((NHV + 3) & ~3, 4);
    // per-vertex entry lists (second pass) to size the CSR
    std::vector<std::vector<uint16_t>> pent(np);
    std::vector<std::vector<uint16_t>> pptr(np);
    std::vector<int> hslot(nc, 0);
    size_t NE = 0;
    std::fill(vstamp.begin(), vstamp.end(), -1);
    for (long long p = 0; p < np; ++p) {
        const long long c0 = p * TB_P, c1 = std::min(no, c0 + TB_P);
        for (size_t k = 0; k < phalo[p].size(); ++k) hslot[phalo[p][k]] = TB_P + (int)k;
        auto slot_of = [&](long long c) -> int { return (c >= c0 && c < c1) ? (int)(c - c0) : hslot[c]; };
        auto &ent = pent[p];
        auto &vp = pptr[p];
        vp.push_back(0);
        for (size_t k = 0; k < pverts[p].size(); ++k) {
            const int v = pverts[p][k];
            for (long long kk = ptr[v]; kk < ptr[v + 1]; ++kk) {
                const long long c = idx[kk];
                const int sl = slot_of(c);
                ent.push_back((uint16_t)sl);
                const unsigned char m = bmask(c);
                if (m)
                    for (int f = 0; f < 3; ++f)
                        if ((m & (1 << f)) && (tvert(c, (f + 1) % 3) == v || tvert(c, (f + 2) % 3) == v))
                            ent.push_back((uint16_t)(0x8000 | (sl << 2) | f));
            }
            if (ent.size() > 0xffff) return fail(ctx, TB_ERR_UNSUPPORTED, "patch vertex adjacency too large");
            vp.push_back((uint16_t)ent.size());
        }
        NE = std::max(NE, ent.size());
    }
    NE = (NE + 7) & ~(size_t)7;
    const size_t off_ctv = (size_t)NHV * sizeof(int32_t);
    const size_t off_vptr = off_ctv + (size_t)TB_P * 3 * sizeof(uint16_t);
    const size_t off_vidx = off_vptr + (((size_t)(NVT + 1) * sizeof(uint16_t) + 3) & ~(size_t)3);
    const size_t stride = (off_vidx + NE * sizeof(uint16_t) + 15) & ~(size_t)15;
    std::vector<unsigned char> tab((size_t)np * stride, 0);
    std::vector<int32_t> nhv((size_t)np * 2, 0);
    std::fill(vstamp.begin(), vstamp.end(), -1);
    for (long long p = 0; p < np; ++p) {
        const long long c0 = p * TB_P, c1 = std::min(no, c0 + TB_P);
        for (size_t k = 0; k < pverts[p].size(); ++k) {
            vstamp[pverts[p][k]] = (int)p;
            vloc[pverts[p][k]] = (int)k;
        }
        unsigned char *blk = tab.data() + (size_t)p * stride;
        int32_t *hids = reinterpret_cast<int32_t *>(blk);
        uint16_t *ctv = reinterpret_cast<uint16_t *>(blk + off_ctv);
        nhv[2 * p] = (int32_t)phalo[p].size();
        nhv[2 * p + 1] = (int32_t)pverts[p].size();
        for (size_t k = 0; k < phalo[p].size(); ++k) hids[k] = (int32_t)dev_cell(phalo[p][k]);
        for (long long c = c0; c < c1; ++c)
            for (int a = 0; a < 3; ++a) ctv[(c - c0) * 3 + a] = (uint16_t)vloc[tvert(c, a)];
        memcpy(blk + off_vptr, pptr[p].data(), pptr[p].size() * sizeof(uint16_t));
        memcpy(blk + off_vidx, pent[p].data(), pent[p].size() * sizeof(uint16_t));
    }
    cudaFree(ctx->d_lim_tab);
    cudaFree(ctx->d_lim_nhv);
    ctx->d_lim_tab = nullptr;
    ctx->d_lim_nhv = nullptr;
    CK(cudaMalloc(&ctx->d_lim_tab, std::max<size_t>(tab.size(), 16)));
    CK(cudaMemcpy(ctx->d_lim_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ctx->d_lim_nhv, sizeof(int32_t) * 2 * np));
    CK(cudaMemcpy(ctx->d_lim_nhv, nhv.data(), sizeof(int32_t) * 2 * np, cudaMemcpyHostToDevice));
    ctx->lim.n_owned = no;
    ctx->lim.n_cells = nc;
    ctx->lim.tab = ctx->d_lim_tab;
    ctx->lim.stride = (long long)stride;
    ctx->lim.NHV = NHV;
    ctx->lim.NVT = NVT;
    ctx->lim.off_ctv = (int)off_ctv;
    ctx->lim.off_vptr = (int)off_vptr;
    ctx->lim.off_vidx = (int)off_vidx;
    ctx->lim.counts = ctx->d_lim_nhv;
    ctx->lim_ready = true;
    return TB_OK;
}

static int limiter_impl(tb_ctx *ctx, const double *c_in, double *c_out, const unsigned long long *push_dst, void *stream) {
    TbRange range(push_dst ? "tb_limiter_apply_fused" : "tb_limiter_apply");
    if (!ctx || !c_in || !c_out) return fail(ctx, TB_ERR_ARG, "null pointer");
    if (c_in == c_out) return fail(ctx, TB_ERR_ARG, "c_out must not alias c_in (use tb_limiter_apply)");
    if (!ctx->lim_ready) {
        int rc = limiter_setup(ctx);
        if (rc != TB_OK) return rc;
    }
    TbLimiterData d = ctx->lim;
    if (push_dst) {
        if (!ctx->fused_ready) return fail(ctx, TB_ERR_STATE, "tb_halo_fused_setup has not been called");
        d.patch_list = ctx->d_fused_order;
        d.halo = ctx->d_fused;
        d.push_dst = push_dst;
        d.n_bpatch = ctx->fused_n_bpatch;
    }
    CK(tb_launch_limiter(d, c_in, c_out, (cudaStream_t)stream));
    ctx->launches += 1;
    return TB_OK;
}

extern "C" int tb_limiter_apply_to(tb_ctx *ctx, const double *c_in, double *c_out, void *stream) {
    return limiter_impl(ctx, c_in, c_out, nullptr, stream);
}

extern "C" int tb_limiter_apply_to_fused(tb_ctx *ctx, const double *c_in, double *c_out, const uint64_t *push_dst,
                                         void *stream) {
    if (!push_dst) return fail(ctx, TB_ERR_ARG, "null push table");
    return limiter_impl(ctx, c_in, c_out, reinterpret_cast<const unsigned long long *>(push_dst), stream);
}

extern "C" int tb_limiter_apply(tb_ctx *ctx, double *c, void *stream) {
    // in place = out of place into a scratch array + copy back of the owned cells (neighbouring patches read the
    // original values of each other's cells); integrators that can swap buffers use tb_limiter_apply_to
    if (!ctx || !c) return fail(ctx, TB_ERR_ARG, "null pointer");
    if (!ctx->d_lim_tmp) CK(cudaMalloc(&ctx->d_lim_tmp, sizeof(double) * (size_t)tb_tracer_len(ctx)));
    int rc = tb_limiter_apply_to(ctx, c, ctx->d_lim_tmp, stream);
    if (rc != TB_OK) return rc;
    CK(cudaMemcpyAsync(c, ctx->d_lim_tmp, sizeof(double) * (size_t)ctx->n_owned * 3, cudaMemcpyDeviceToDevice,
                       (cudaStream_t)stream));
    return TB_OK;
}

// ------------------------------------------------------------------ layout conversion & diagnostics
extern "C" int tb_state_from_fields(tb_ctx *ctx, const double *uv, const double *eta, const int32_t *node_map,
                                    double *state, void *stream) {
    if (!ctx || !uv || !eta || !node_map || !state) return fail(ctx, TB_ERR_ARG, "null pointer");
    CK(tb_launch_state_from_fields(uv, eta, node_map, state, ctx->n_owned, (cudaStream_t)stream));
    ctx->launches++;
    return TB_OK;
}
extern "C" int tb_state_to_fields(tb_ctx *ctx, const double *state, const int32_t *node_map, double *uv, double *eta,
                                  void *stream) {
    if (!ctx || !uv || !eta || !node_map || !state) return fail(ctx, TB_ERR_ARG, "null pointer");
    CK(tb_launch_state_to_fields(state, node_map, uv, eta, ctx->n_owned, (cudaStream_t)stream));
    ctx->launches++;
    return TB_OK;
}
extern "C" int tb_tracer_from_field(tb_ctx *ctx, const double *q, const int32_t *node_map, double *c, void *stream) {
    if (!ctx || !q || !node_map || !c) return fail(ctx, TB_ERR_ARG, "null pointer");
    CK(tb_launch_tracer_from_field(q, node_map, c, ctx->n_owned, (cudaStream_t)stream));
    ctx->launches++;
    return TB_OK;
}
extern "C" int tb_tracer_to_field(tb_ctx *ctx, const double *c, const int32_t *node_map, double *q, void *stream) {
    if (!ctx || !q || !node_map || !c) return fail(ctx, TB_ERR_ARG, "null pointer");
    CK(tb_launch_tracer_to_field(c, node_map, q, ctx->n_owned, (cudaStream_t)stream));
    ctx->launches++;
    return TB_OK;
}
extern "C" int tb_swe_integrals(tb_ctx *ctx, const double *state, double *out, void *stream) {
    if (!ctx || !state || !out) return fail(ctx, TB_ERR_ARG, "null pointer");
    {
        int rc = sync_fields(ctx, (cudaStream_t)stream);
        if (rc != TB_OK) return rc;
    }
    CK(tb_launch_swe_integrals(state, ctx->d_area, ctx->d_bath3, ctx->n_owned, ctx->d_partial, out, (cudaStream_t)stream));
    ctx->launches += 2;
    return TB_OK;
}
extern "C" int tb_stage_integrals(tb_ctx *ctx, int enable) {
    if (!ctx) return TB_ERR_ARG;
    ctx->stage_integrals = enable != 0;
    return TB_OK;
}
extern "C" int tb_stage_integrals_finish(tb_ctx *ctx, double *out, void *stream) {
    if (!ctx || !out) return fail(ctx, TB_ERR_ARG, "null pointer");
    if (!ctx->d_stage_partial) return fail(ctx, TB_ERR_STATE, "no stage launch has produced fused diagnostics yet");
    CK(tb_launch_patch_partials_final(ctx->d_stage_partial, ctx->n_patches, out, (cudaStream_t)stream));
    ctx->launches += 1;
    return TB_OK;
}
extern "C" int tb_tracer_integrals(tb_ctx *ctx, const double *c, const double *swe_state, double *out, void *stream) {
    if (!ctx || !c || !out) return fail(ctx, TB_ERR_ARG, "null pointer");
    if (ctx->nonlinear && !swe_state) return fail(ctx, TB_ERR_ARG, "swe_state required for the nonlinear total depth");
    {
        int rc = sync_fields(ctx, (cudaStream_t)stream);
        if (rc != TB_OK) return rc;
    }
    CK(tb_launch_tracer_integrals(c, swe_state, ctx->d_area, ctx->d_bath3, ctx->n_owned, ctx->nonlinear,
                                  ctx->wd_on && ctx->nonlinear, ctx->wd_alpha * ctx->wd_alpha, ctx->nquad, ctx->d_partial,
                                  out, (cudaStream_t)stream));
    ctx->launches += 2;
    return TB_OK;
}
extern "C" int tb_lincomb(tb_ctx *ctx, int n, const double *const *x, const double *w, double *out, int64_t len,
                          void *stream) {
    if (!ctx || !x || !w || !out || n < 1 || n > 6 || len < 0)
        return fail(ctx, TB_ERR_ARG, "bad linear-combination arguments");
    if (reinterpret_cast<uintptr_t>(out) & 15) return fail(ctx, TB_ERR_ARG, "operands must be 16-byte aligned");
    for (int j = 0; j < n; ++j)
        if (!x[j] || (reinterpret_cast<uintptr_t>(x[j]) & 15)) return fail(ctx, TB_ERR_ARG, "null or misaligned operand");
    CK(tb_launch_lincomb(n, x, w, out, len, (cudaStream_t)stream));
    ctx->launches += len > 0;
    return TB_OK;
}
extern "C" int tb_gather_cells(tb_ctx *ctx, const double *state, const int32_t *idx, int64_t n, int rec_len,
                               double *buf, void *stream) {
    if (!ctx || (n > 0 && (!state || !idx || !buf))) return fail(ctx, TB_ERR_ARG, "null pointer");
    CK(tb_launch_gather_cells(state, idx, n, rec_len, buf, (cudaStream_t)stream));
    ctx->launches += n > 0;
    return TB_OK;
}
extern "C" int tb_scatter_cells(tb_ctx *ctx, const double *buf, const int32_t *idx, int64_t n, int rec_len,
                                double *state, void *stream) {
    if (!ctx || (n > 0 && (!state || !idx || !buf))) return fail(ctx, TB_ERR_ARG, "null pointer");
    CK(tb_launch_scatter_cells(buf, idx, n, rec_len, state, (cudaStream_t)stream));
    ctx->launches += n > 0;
    return TB_OK;
}
extern "C" int tb_push_cells(tb_ctx *ctx, const double *state, const int32_t *idx, const uint64_t *dst_ptrs, int64_t n,
                             int rec_len, void *stream) {
    TbRange range("tb_push_cells");
    if (!ctx || (n > 0 && (!state || !idx || !dst_ptrs))) return fail(ctx, TB_ERR_ARG, "null pointer");
    CK(tb_launch_push_cells(state, idx, reinterpret_cast<const unsigned long long *>(dst_ptrs), n, rec_len,
                            (cudaStream_t)stream));
    ctx->launches += n > 0;
    return TB_OK;
}
extern "C" int tb_set_patch_range(tb_ctx *ctx, int64_t first, int64_t count) {
    if (!ctx) return TB_ERR_ARG;
    ctx->range_first = first;
    ctx->range_count = count;
    ctx->patch_list = nullptr;
    ctx->patch_list_n = 0;
    return TB_OK;
}
extern "C" int tb_set_patch_list(tb_ctx *ctx, const int32_t *list, int64_t n) {
    if (!ctx || (n > 0 && !list) || n > ctx->n_patches) return fail(ctx, TB_ERR_ARG, "bad patch list");
    ctx->patch_list = n > 0 ? list : nullptr;
    ctx->patch_list_n = n > 0 ? n : 0;
    return TB_OK;
}

extern "C" int tb_selftest_math(tb_ctx *ctx, const double *x, double *out, int64_t n, void *stream) {
    if (!ctx || !x || !out) return fail(ctx, TB_ERR_ARG, "null pointer");
    CK(tb_launch_test_math(x, out, (int)n, (cudaStream_t)stream));
    return TB_OK;
}

extern "C" int tb_set_cell_quadrature(tb_ctx *ctx, int n, const double *lam, const double *w) {
    if (!ctx || !lam || !w || n < 1 || n > TB_MAX_QUAD) return fail(ctx, TB_ERR_ARG, "bad quadrature rule");
    CK(cudaDeviceSynchronize());
    CK(tb_set_quadrature(n, lam, w));
    CK(tb_set_quadrature_tracer(n, lam, w));
    ctx->nquad = n;
    return TB_OK;
}
