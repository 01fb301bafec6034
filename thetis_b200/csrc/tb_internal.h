// Internal structures shared by the kernels and the C-ABI glue (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/thetis_b200.h"

#ifndef TB_P
#define TB_P 128            // cells per patch == threads per CTA (one thread per cell)
#endif
#define TB_MAX_SLOTS 16     // distinct boundary markers
#define TB_MAX_QUAD 12      // max cell quadrature points
#define TB_MAX_BANKS 4      // banks of Function-valued boundary data (one per RK stage of a graph-replayed step)

// coefficient descriptor: mode 0 = None, 1 = Constant, 2 = P1 vertex column(s) of the static block,
// 3 = discontinuous P1DG field stored per cell node: cell[(cell*3 + node)*nc + comp] (generic kernels only)
struct TbCoef {
    int mode;
    int col;
    double v0, v1;
    const double *cell;
    int nc, pad_;
};

struct TbBcSlot {
    int marker;
    int opcode;      // OR of TB_BC_*
    int arr_mask;    // tags whose datum is a per-facet-node array instead of a constant
    int pad;
    double elev, uvx, uvy, un, flux, value;   // value: tracer 'value'; in a shallow-water slot the 'drag' coefficient
    double bnd_len;
    double diff_flux;    // tracer 'diff_flux' datum (tracer_eq_2d.py:264-265)
};

struct TbBcTable {
    int n_slots;
    int pad;
    const int *bf_slot;          // [n_bfacets] slot of each exterior facet
    const int *bf_row;           // [n_bfacets] row of the facet in the compact per-tag arrays (grouped by slot)
    const double *ext_elev;      // [n_bfacets*2], indexed by bf_row
    const double *ext_uv;        // [n_bfacets*4]
    const double *ext_un;        // [n_bfacets*2]
    const double *ext_flux;      // [n_bfacets*2]
    const double *ext_value;     // [n_bfacets*2]
    TbBcSlot slots[TB_MAX_SLOTS];
};

// static per-patch block (one TMA bulk copy):
//   double col[ncol][NV]   vertex columns: 0 = x, 1 = y, 2 = bathymetry, then optional fields
//   uint16 cv[TB_P][3]     local vertex ids
//   int32  cn[TB_P][3]     >= 0: (smem cell index)*4 + local facet in neighbour;  < 0: -(1 + exterior facet)
//   uint16 hcv[NH][3]      only when a SIPG term (viscosity / diffusion) is active: local vertex ids of the halo
//                          cells (their vertices are then part of the patch's vertex columns), off_hcv >= 0
struct TbPatchLayout {
    const unsigned char *sblk;
    long long stride;
    int NV, NH, ncol;
    int off_cv, off_cn;
    int off_hcv, pad_;
    const int *halo_ids;         // [n_patches*NH] cell ids of off-patch facet neighbours
    const int *halo_cnt;         // [n_patches]
};

// Fused halo exchange of the distributed SWE stage (tb_swe_stage_fused): the CTAs of the first n_bpatch patches of
// the launch order hold every cell a peer rank needs.  They (1) wait until every peer they receive from has published
// the ghost records of the stage input (per-peer flag >= epoch), (2) evaluate their patch, (3) store the records the
// peers need straight into the peers' ghost blocks (NVLink peer stores) and (4) the last of them to finish publishes
// epoch + 1 in every receiving peer's flag array.  Device-resident so that a captured CUDA graph replays correctly.
#define TB_MAX_PEERS 16
#define TB_HAVE_HALO_FUSED 1
struct TbHaloFused {
    unsigned long long *epoch;         // [1] fused stage launches completed by this rank
    unsigned int *done_count;          // [1] boundary CTAs of the running launch that finished their push
    int *error;                        // [1] set when a flag wait timed out
    const unsigned long long *flags;   // [world] local: flags[q] = epochs published by peer q
    const int *push_ptr;               // [n_bpatch+1] CSR over the boundary CTAs of the launch order
    const int *push_cell;              // [n_entries] cell index inside the patch
    int n_recv, n_send;
    int recv_peer[TB_MAX_PEERS];                   // ranks whose flags this rank waits for
    unsigned long long *remote_flag[TB_MAX_PEERS]; // &flags[my rank] on every rank this rank sends to
};

struct TbSweParams {
    const double *u_in;
    const double *u0;
    double *u_out;
    TbPatchLayout pl;
    const int *patch_list;       // optional: patch of CTA b is patch_list[b] (boundary / interior split), else patch_first + b
    int n_owned;
    int patch_first;
    double a0, a1, bdt;
    double g, rho0, lf_sigma, eps2, wd_alpha2;
    int lf_on, wd_on, use_quad, nquad;
    int force_generic, adv_on;    // developer switch: always run the generic (SPEC 0) kernel; momentum advection on/off
    double *partials;             // optional [n_patches][4]: fused diagnostics of u_out (int eta^2, |u|^2, eta, eta+b)
    const TbHaloFused *halo;      // optional (device): fused halo exchange, CTAs [0, n_bpatch) push
    const unsigned long long *push_dst;   // [n_entries] peer addresses of the pushed records for THIS output buffer
    int n_bpatch;
    int wd_mass;                  // TB_OPT_WD_DISPLACED_MASS: displaced-mass elevation update (tb_wd_mass.cuh)
    TbCoef cor, man, cd, lin, wind, pa, msrc, vsrc, visc, nik, wda;
    double kappa;                 // physical_constants['von_karman'] (Nikuradse drag)
    double sipg;                  // sipg_factor (HorizontalViscosityTerm, shallowwater_eq.py:558)
    int graddiv, graddepth;       // use_grad_div_viscosity_term, use_grad_depth_viscosity_term
    TbBcTable bc;
};

struct TbTracerParams {
    const double *c_in;
    const double *c0;
    double *c_out;
    const double *swe;           // frozen SWE state (cell records)
    TbPatchLayout pl;
    int n_owned;
    int patch_first;
    double a0, a1, bdt;
    double corr, lf_sigma;
    int lf_on, nonlin, wd_on, pad;
    double wd_alpha2;
    TbCoef src, diff;
    double sipg;                  // sipg_factor_tracer
    int conservative, nquad;
    int force_generic, pad1;
    const int *patch_list;        // optional launch order (fused halo exchange: partition-boundary patches first)
    const TbHaloFused *halo;      // optional (device): CTAs [0, n_bpatch) wait for / push ghost records
    const unsigned long long *push_dst;   // [n_entries] peer addresses of the pushed records for THIS output buffer
    int n_bpatch, pad2;
    TbBcTable bc;
};

// kernel launchers (tb_kernels.cu)
cudaError_t tb_launch_swe_stage(const TbSweParams &p, bool nonlinear, int n_patches, size_t smem, cudaStream_t s);
cudaError_t tb_launch_tracer_stage(const TbTracerParams &p, int n_patches, size_t smem, cudaStream_t s);
cudaError_t tb_set_quadrature(int n, const double *lam, const double *w);
cudaError_t tb_set_quadrature_tracer(int n, const double *lam, const double *w);
cudaError_t tb_launch_lincomb(int n, const double *const *x, const double *w, double *out, long long len, cudaStream_t s);
cudaError_t tb_launch_tracer_integrals(const double *c, const double *swe, const double *area, const double *bath3,
                                       long long n_owned, int nonlin, int wd_on, double alpha2, int nquad,
                                       double *partial, double *out, cudaStream_t s);
int tb_swe_stage_spec(const TbSweParams &p, bool nonlinear);
size_t tb_swe_smem_bytes(const TbPatchLayout &pl);
size_t tb_tracer_smem_bytes(const TbPatchLayout &pl);
cudaError_t tb_kernels_init();          // per device: raises the dynamic shared-memory limit of the stage kernels
cudaError_t tb_tracer_kernels_init();
cudaError_t tb_launch_test_math(const double *x, double *out, int n, cudaStream_t s);

cudaError_t tb_launch_state_from_fields(const double *uv, const double *eta, const int32_t *node_map,
                                        double *state, long long n_cells, cudaStream_t s);
cudaError_t tb_launch_state_to_fields(const double *state, const int32_t *node_map, double *uv, double *eta,
                                      long long n_cells, cudaStream_t s);
cudaError_t tb_launch_tracer_from_field(const double *q, const int32_t *node_map, double *c, long long n_cells,
                                        cudaStream_t s);
cudaError_t tb_launch_tracer_to_field(const double *c, const int32_t *node_map, double *q, long long n_cells,
                                      cudaStream_t s);
cudaError_t tb_launch_gather_cells(const double *state, const int32_t *idx, long long n, int rec, double *buf,
                                   cudaStream_t s);
cudaError_t tb_launch_scatter_cells(const double *buf, const int32_t *idx, long long n, int rec, double *state,
                                    cudaStream_t s);
cudaError_t tb_launch_push_cells(const double *state, const int32_t *idx, const unsigned long long *dst, long long n,
                                 int rec, cudaStream_t s);
cudaError_t tb_launch_halo_fused_wait(const TbHaloFused *hf, cudaStream_t s);
cudaError_t tb_launch_update_columns(unsigned char *sblk, long long stride, int NV, long long n_patches,
                                     const int32_t *patch_vglob, const double *vert, int col, int ncomp, cudaStream_t s);
cudaError_t tb_launch_patch_partials_final(const double *partial, long long n, double *out, cudaStream_t s);
cudaError_t tb_launch_swe_integrals(const double *state, const double *area, const double *bath3, long long n_owned,
                                    double *partial, double *out, cudaStream_t s);
#define TB_NRED 296           // CTAs of the two-pass reductions (2 per SM)
// limiter: one patch-staged kernel (tb_tracer.cu); per patch a static table at tab + patch*stride:
//   int32  hids[NHV]      device cell ids of the vertex halo (cells outside the patch sharing a vertex with it)
//   uint16 ctv[TB_P][3]   patch-local topological vertex of each own-cell node
//   uint16 vptr[NVT+1]    CSR over the patch vertices into vidx
//   uint16 vidx[NE]       cells around the vertex as shared-memory slots (own cell t = t, halo cell h = TB_P + h), or
//                         0x8000 | slot << 2 | f : exterior facet f of that cell touches the vertex
struct TbLimiterData {
    long long n_owned, n_cells;
    const unsigned char *tab;
    long long stride;
    int NHV, NVT;                // padded sizes: vertex-halo cells / topological vertices per patch
    int off_ctv, off_vptr, off_vidx, pad_;
    const int *counts;           // [n_patches][2] (vertex-halo cells, vertices) of each patch
    const int *patch_list;       // optional launch order (fused halo exchange: partition-boundary patches first)
    const TbHaloFused *halo;     // optional (device): CTAs [0, n_bpatch) wait for / push ghost records
    const unsigned long long *push_dst;
    int n_bpatch, pad2_;
};
cudaError_t tb_launch_limiter(const TbLimiterData &d, const double *c_in, double *c_out, cudaStream_t s);
