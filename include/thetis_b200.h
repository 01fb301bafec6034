/*
 * thetis_b200 -- C ABI of the B200-native explicit P1DG shallow-water stepper.
 *
 * This is the drop-in boundary for ONE hot path of thetisproject/thetis: the
 * explicit SSPRK33 step of `ShallowWaterEquations` on P1DG-P1DG triangles, the
 * explicit 2-D tracer advection and `VertexBasedP1DGLimiter`.  The reference has
 * no FFI of its own (it is pure Python on Firedrake); each entry point below
 * names the reference interface whose work it replaces (paths relative to the
 * thetis repository).  The Python class that mirrors
 * `thetis.timeintegrator.TimeIntegrator` binds these with ctypes
 * (thetis_b200/_lib.py); INTEGRATION.md shows the stub a Thetis maintainer adds.
 *
 * Conventions: extern "C"; every function returns 0 on success or a negative
 * tb_status; no exceptions, no ownership transfer.  Mesh/coefficient arrays
 * handed to tb_create / tb_set_* are HOST pointers read once at set-up.  All
 * bulk state pointers (u_in, u_out, ...) are caller-owned DEVICE pointers
 * (e.g. torch.Tensor.data_ptr()).  `stream` is a cudaStream_t passed as void*;
 * nothing synchronises the stream unless stated.
 *
 * Device state layout ("cell records"): for local cell c, 9 doubles
 *   [u0x u0y u1x u1y u2x u2y eta0 eta1 eta2]      (nodes in the cell's CCW order)
 * in an array of tb_state_len() doubles: owned cells padded to a multiple of
 * the patch size, followed by ghost cells (multi-GPU halo).
 */
#ifndef THETIS_B200_H
#define THETIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb_ctx tb_ctx;

typedef enum {
    TB_OK = 0,
    TB_ERR_ARG = -1,        /* invalid argument */
    TB_ERR_CUDA = -2,       /* CUDA runtime error (see tb_last_error) */
    TB_ERR_UNSUPPORTED = -3,/* configuration outside the accelerated path */
    TB_ERR_STATE = -4       /* call order / missing data */
} tb_status;

/* Mesh connectivity, HOST arrays, read once (what the adaptor extracts from
 * Firedrake's mesh.coordinates / cell_node_map / exterior_facets; see
 * SURVEY.md 8b).  Cells must be counter-clockwise. */
typedef struct {
    int64_t n_cells;        /* owned + ghost cells                               */
    int64_t n_owned;        /* cells [0, n_owned) are advanced; the rest are ghosts */
    int64_t n_vertices;     /* geometric vertices                                */
    int64_t n_bfacets;      /* exterior facets                                   */
    const double  *coords;  /* [n_vertices*2]                                    */
    const int32_t *cells;   /* [n_cells*3] geometric vertex ids                  */
    const int32_t *nbr;     /* [n_cells*3] neighbour across local facet i (opposite
                               vertex i), or -(1+k) for exterior facet k; for
                               ghost cells entries may be INT32_MIN (unknown)    */
    const int8_t  *nbr_lf;  /* [n_cells*3] local facet number inside the neighbour */
    const int32_t *bf_marker;/* [n_bfacets] boundary marker id                    */
    const int32_t *topo;    /* [n_vertices] topological vertex id (periodic meshes;
                               may be NULL = identity); used by the limiter      */
} tb_mesh;

/* Scalar options: `ModelOptions2d` flags consumed on this path
 * (thetis/options.py:596-611,712,871-881) and thetis/physical_constants.py. */
typedef enum {
    TB_OPT_G_GRAV = 0,                 /* physical_constants['g_grav']            */
    TB_OPT_RHO0 = 1,                   /* physical_constants['rho0']              */
    TB_OPT_NONLINEAR = 2,              /* use_nonlinear_equations                 */
    TB_OPT_LAX_FRIEDRICHS = 3,         /* use_lax_friedrichs_velocity             */
    TB_OPT_LF_SCALING = 4,             /* lax_friedrichs_velocity_scaling_factor  */
    TB_OPT_NORM_SMOOTHER = 5,          /* norm_smoother                           */
    TB_OPT_WETTING_DRYING = 6,         /* use_wetting_and_drying                  */
    TB_OPT_WD_ALPHA = 7,               /* wetting_and_drying_alpha (constant)     */
    TB_OPT_LF_TRACER = 8,              /* use_lax_friedrichs_tracer               */
    TB_OPT_LF_TRACER_SCALING = 9,      /* lax_friedrichs_tracer_scaling_factor    */
    TB_OPT_TRACER_VEL_FACTOR = 10,     /* tracer_advective_velocity_factor        */
    TB_OPT_FORCE_GENERIC_KERNEL = 11,  /* developer/test switch: bypass the specialised stage kernels */
    TB_OPT_SIPG_FACTOR = 12,           /* sipg_factor (options.py:730; shallowwater_eq.py:558)      */
    TB_OPT_SIPG_FACTOR_TRACER = 13,    /* sipg_factor_tracer (options.py:732; tracer_eq_2d.py:235)  */
    TB_OPT_GRAD_DIV_VISCOSITY = 14,    /* use_grad_div_viscosity_term (options.py:597)              */
    TB_OPT_GRAD_DEPTH_VISCOSITY = 15,  /* use_grad_depth_viscosity_term (options.py:602), default on */
    TB_OPT_TRACER_CONSERVATIVE = 16,   /* tracer use_conservative_form (options.py:543; tracer_eq_2d.py:323-437) */
    TB_OPT_MOMENTUM_ADVECTION = 17,    /* 0: no HorizontalAdvectionTerm although the depth is nonlinear
                                          (ModeSplit2DEquations, shallowwater_eq.py:931-966); default 1 */
    TB_OPT_VON_KARMAN = 18,            /* physical_constants['von_karman'] (Nikuradse drag, shallowwater_eq.py:696) */
    TB_OPT_WD_DISPLACED_MASS = 19      /* wetting-drying, nonlinear equations: the elevation update of every Shu-Osher
                                          stage (a0 + a1 = 1) advances the reference's own mass functional
                                          int (eta + f(b + eta)) phi (shallowwater_eq.py:917-920, :834-850) instead of the
                                          plain mass: cell-local Newton solve in the stage kernel's epilogue (generic
                                          kernel).  Default 0 = plain mass.  DESIGN.md section 6. */
} tb_option;

/* Coefficient fields: the `fields` dict of solver2d.py:546-558 plus bathymetry. */
typedef enum {
    TB_F_BATHYMETRY = 0,       /* DepthExpression.bathymetry_2d (utility.py:965)  */
    TB_F_CORIOLIS = 1,         /* 'coriolis'                                      */
    TB_F_MANNING = 2,          /* 'manning_drag_coefficient'                      */
    TB_F_QUAD_DRAG = 3,        /* 'quadratic_drag_coefficient'                    */
    TB_F_LINEAR_DRAG = 4,      /* 'linear_drag_coefficient'                       */
    TB_F_WIND_STRESS = 5,      /* 'wind_stress' (2 components)                    */
    TB_F_ATM_PRESSURE = 6,     /* 'atmospheric_pressure'                          */
    TB_F_MOMENTUM_SOURCE = 7,  /* 'momentum_source' (2 components)                */
    TB_F_VOLUME_SOURCE = 8,    /* 'volume_source'                                 */
    TB_F_TRACER_SOURCE = 9,    /* 'source-<label>' of the tracer equation         */
    TB_F_VISCOSITY = 10,       /* 'viscosity_h' (HorizontalViscosityTerm, shallowwater_eq.py:554-616) */
    TB_F_DIFFUSIVITY = 11,     /* 'diffusivity_h-<label>' (HorizontalDiffusionTerm, tracer_eq_2d.py:226-278) */
    TB_F_NIKURADSE = 12,       /* 'nikuradse_bed_roughness' (QuadraticDragTerm, shallowwater_eq.py:689-697) */
    TB_F_WD_ALPHA = 13,        /* wetting_and_drying_alpha as a P1 Function (solver2d.py:279-287, utility.py:981-983);
                                  overrides TB_OPT_WD_ALPHA in the shallow-water stage */
    TB_F_COUNT = 14
} tb_field;

/* Boundary tags (shallowwater_eq.py:243-267); a marker's opcode is the OR of
 * the tags present.  0 = closed (land) boundary. */
enum { TB_BC_ELEV = 1, TB_BC_UV = 2, TB_BC_UN = 4, TB_BC_FLUX = 8, TB_BC_VALUE = 16,
       TB_BC_DIFF_FLUX = 64, /* tracer 'diff_flux' (tracer_eq_2d.py:264-265); 32 is reserved */
       TB_BC_DRAG = 128      /* shallow water 'drag': quadratic friction of the tangential velocity on the marker's
                                facets (BoundaryDragTerm, shallowwater_eq.py:704-726); not an open-boundary tag, it
                                combines with a closed boundary or with any of the open ones (:286-296) */ };

/* ---- lifetime ---------------------------------------------------------- */
int tb_create(tb_ctx **out, const tb_mesh *mesh, int device);
int tb_destroy(tb_ctx *ctx);
const char *tb_last_error(const tb_ctx *ctx);   /* ctx may be NULL: create errors */
int tb_version(void);

/* ---- sizes ------------------------------------------------------------- */
int64_t tb_state_len(const tb_ctx *ctx);        /* doubles in an SWE state array    */
int64_t tb_tracer_len(const tb_ctx *ctx);       /* doubles in a tracer array (3/cell) */
int64_t tb_patch_size(const tb_ctx *ctx);
int64_t tb_n_patches(const tb_ctx *ctx);

/* ---- options / coefficients / boundary conditions ---------------------- */
int tb_set_option(tb_ctx *ctx, int option, double value);
/* constant coefficient (Firedrake `Constant`); ncomp = 1 or 2 */
int tb_set_field_const(tb_ctx *ctx, int field, const double *value, int ncomp);
/* P1 coefficient given at the geometric vertices, HOST [n_vertices*ncomp] */
int tb_set_field_vertex(tb_ctx *ctx, int field, const double *values, int ncomp);
/* Genuinely discontinuous P1DG coefficient (e.g. Coriolis / sources projected into H_2d or U_2d,
 * test/swe2d/test_steady_state_basin_mms.py:169-177; P1DG atmospheric pressure, test_atmospheric_pressure.py:57-63):
 * HOST [n_owned*3*ncomp], values at the CCW nodes of every owned cell in the context's cell order.  Copied
 * asynchronously on `stream` (cheap to repeat every RK stage).  Only for coefficients of cell terms: bathymetry,
 * viscosity, diffusivity and the wetting-drying alpha enter facet terms and must be continuous (TB_ERR_UNSUPPORTED). */
int tb_set_field_cell(tb_ctx *ctx, int field, const double *values, int ncomp, void *stream);
int tb_clear_field(tb_ctx *ctx, int field);     /* field = None                     */
/* Apply pending coefficient changes stream-ordered on `stream` (every stage launch does this itself; call it
 * explicitly before replaying a captured CUDA graph).  New VALUES of an existing P1 field rewrite only that field's
 * columns (async copy + one scatter kernel, no device synchronisation: time-dependent wind / pressure forcing in
 * update_forcings); structural changes rebuild the per-patch blocks (synchronising, set-up time only). */
int tb_sync_fields(tb_ctx *ctx, void *stream);
/* Boundary condition of one marker for equation eq (0 = shallow water,
 * 1 = tracer): opcode = OR of TB_BC_*, consts = {elev, uv_x, uv_y, un, flux, value, diff_flux, drag}.
 * replaces ShallowWaterTerm.get_bnd_functions (shallowwater_eq.py:232-272)
 * and TracerTerm.get_bnd_functions (tracer_eq_2d.py:78-115) */
int tb_set_bc(tb_ctx *ctx, int eq, int marker, int opcode, const double consts[8]);
/* The marker has no entry in this equation's bnd_conditions dict (closed boundary; each tracer equation is built
 * from its own dict, solver2d.py:580-598, while the device context is shared). */
int tb_clear_bc(tb_ctx *ctx, int eq, int marker);
/* Spatially varying datum for one tag of one marker: HOST values at the two
 * nodes of every exterior facet of the mesh, [n_bfacets*2*ncomp] (entries of
 * other markers ignored).  Copied asynchronously on `stream`. */
int tb_set_bc_array(tb_ctx *ctx, int eq, int marker, int tag, const double *values,
                    int ncomp, void *stream);
int tb_set_boundary_length(tb_ctx *ctx, int marker, double length); /* utility.py:821-832 */
/* Banks of the Function-valued shallow-water boundary data (0 <= bank < 4): subsequent tb_set_bc_array calls fill,
 * and subsequent stage launches read, the selected bank.  Lets a whole RK step be captured in ONE CUDA graph whose
 * stage i reads bank i, while `update_forcings(t + c_i dt)` (rungekutta.py:933-934) still provides different boundary
 * data to every stage: the host fills the banks, then replays the graph.  Default bank 0; a new bank starts as a
 * copy of bank 0. */
int tb_set_bc_bank(tb_ctx *ctx, int bank);

/* Cell quadrature for the non-polynomial cell integrands (Manning drag, wind
 * stress / H, wetting-drying): n <= 12 points, barycentric lam[n*3], weights
 * summing to 1.  Default: the 6-point degree-3 Strang-Fix rule (the reference
 * asks for degree 2p+1 = 3, shallowwater_eq.py:225-230; FIAT's default rule for
 * that degree is version dependent).  The rule is held in
 * constant memory of the device, i.e. it is shared by every context of the
 * process, and tb_create resets it to the default.  The specialised stage
 * kernels exploit the structure of the default rule (one symmetric orbit,
 * equal weights); any other rule is served by the generic stage kernel. */
int tb_set_cell_quadrature(tb_ctx *ctx, int n, const double *lam, const double *w);

/* ---- the hot path ------------------------------------------------------ */
/* One Shu-Osher stage of ERKGenericShuOsher.solve_stage (rungekutta.py:929-946)
 * fused with the residual assembly (shallowwater_eq.py:886-890), the P1DG mass
 * solve (rungekutta.py:921-924) and the stage update:
 *      u_out = a0*u0 + a1*u_in + b_dt * M^-1 R(u_in)
 * u0 may be NULL when a0 == 0.  u_out must not alias u_in. */
int tb_swe_stage(tb_ctx *ctx, double a0, double a1, double b_dt,
                 const double *u_in, const double *u0, double *u_out, void *stream);
/* Parity hook: k = M^-1 R(u)  (the `tendency` of rungekutta.py:940 for dt = 1) */
int tb_swe_tendency(tb_ctx *ctx, const double *u, double *k_out, void *stream);
/* Same for TracerEquation2D (tracer_eq_2d.py:147-193, 293-298) with the frozen
 * SWE state `swe_state` providing uv_2d / elev_2d (solver2d.py:580-598). */
int tb_tracer_stage(tb_ctx *ctx, double a0, double a1, double b_dt,
                    const double *c_in, const double *c0, double *c_out,
                    const double *swe_state, void *stream);
/* VertexBasedP1DGLimiter.apply (limiter.py:182-198), in place. */
int tb_limiter_apply(tb_ctx *ctx, double *c, void *stream);
/* Same, out of place: c_out (owned cells) = limited c_in; ghost cells of c_out are not written.  One patch-staged
 * kernel (bounds in shared memory); the integrator swaps its buffers instead of copying back. */
int tb_limiter_apply_to(tb_ctx *ctx, const double *c_in, double *c_out, void *stream);

/* ---- layout conversion & diagnostics ----------------------------------- */
/* Thetis' mixed Function layout <-> cell records.  uv: [n_nodes*2] interleaved,
 * eta: [n_nodes]; node_map: DEVICE int32 [n_cells*3] giving, for local cell c and
 * CCW local node a, the index of that dof in the Thetis arrays (cell_node_map
 * composed with the renumbering).  All pointers are device pointers. */
int tb_state_from_fields(tb_ctx *ctx, const double *uv, const double *eta,
                         const int32_t *node_map, double *state, void *stream);
int tb_state_to_fields(tb_ctx *ctx, const double *state, const int32_t *node_map,
                       double *uv, double *eta, void *stream);
int tb_tracer_from_field(tb_ctx *ctx, const double *q, const int32_t *node_map,
                         double *c, void *stream);
int tb_tracer_to_field(tb_ctx *ctx, const double *c, const int32_t *node_map,
                       double *q, void *stream);
/* out[0] = int eta^2 dx, out[1] = int |u|^2 dx, out[2] = int eta dx over owned
 * cells (print_state norms, solver2d.py:955-956; VolumeConservation2DCallback).
 * `out` is a DEVICE pointer to 4 doubles. */
int tb_swe_integrals(tb_ctx *ctx, const double *state, double *out, void *stream);
/* Tracer diagnostics reduced on the device (callback.py:369-392,448-484): out[0] = int c dx
 * (ConservativeTracerMassConservation2DCallback), out[1] = int H c dx with H = total depth of `swe_state`
 * (TracerMassConservation2DCallback / comp_tracer_mass_2d), out[2] = min c, out[3] = max c
 * (TracerOvershootCallBack).  `out`: DEVICE pointer to 4 doubles.  tb_swe_integrals' out[3] is
 * int (eta + bathymetry) dx (VolumeConservation2DCallback / comp_volume_2d). */
int tb_tracer_integrals(tb_ctx *ctx, const double *c, const double *swe_state, double *out, void *stream);
/* Fused variant of tb_swe_integrals: while enabled, every tb_swe_stage launch also reduces the four integrals of the
 * state it WRITES (u_out) per patch in its epilogue; tb_stage_integrals_finish sums the per-patch values (fixed
 * order, deterministic) into DEVICE out[4].  Saves the extra pass over the state that print_state / the volume
 * callback would otherwise need after the last RK stage (solver2d.py:955-956). */
int tb_stage_integrals(tb_ctx *ctx, int enable);
int tb_stage_integrals_finish(tb_ctx *ctx, double *out, void *stream);
/* out = sum_j w[j]*x[j], j < n <= 6, over `len` doubles (16-byte aligned operands): the stage combinations of the Butcher-form
 * integrators (ERKGeneric.update_solution / get_final_solution, rungekutta.py:816-852).  x: HOST array of n
 * DEVICE pointers, w: HOST weights.  out may alias any x[j]. */
int tb_lincomb(tb_ctx *ctx, int n, const double *const *x, const double *w, double *out, int64_t len, void *stream);

/* ---- multi-GPU halo (one-deep element halo) -------------- */
/* Gather the records of `n` cells listed in DEVICE idx into a contiguous
 * buffer / scatter them back; used to pack the per-peer send buffers that
 * torch.distributed (NCCL) exchanges once per RK stage. */
int tb_gather_cells(tb_ctx *ctx, const double *state, const int32_t *idx, int64_t n,
                    int rec_len, double *buf, void *stream);
int tb_scatter_cells(tb_ctx *ctx, const double *buf, const int32_t *idx, int64_t n,
                     int rec_len, double *state, void *stream);
/* Peer push over NVLink: record of cell idx[h] is stored to the address dst_ptrs[h] (DEVICE array of device
 * pointers into peer GPUs' ghost blocks, e.g. from torch symmetric memory).  Replaces pack + all-to-all. */
int tb_push_cells(tb_ctx *ctx, const double *state, const int32_t *idx, const uint64_t *dst_ptrs, int64_t n,
                  int rec_len, void *stream);
/* Fused compute + halo exchange (the exchange hidden in every `self.solver.solve()` of the reference,
 * rungekutta.py:940, PETSc SF halo update): ONE launch per RK stage evaluates the partition-boundary patches first,
 * stores the records the peer ranks need straight into their ghost blocks from the kernel epilogue (NVLink peer
 * stores) and publishes a per-peer epoch flag; the boundary patches of the next stage wait for the flags of the ranks
 * they receive from.  No pack kernel, no collective, no host involvement; replayable from a CUDA graph. */
typedef struct {
    int64_t n_bpatch;            /* leading entries of patch_order that hold cells a peer needs        */
    const int32_t *patch_order;  /* HOST [tb_n_patches] launch order, a permutation, boundary patches first */
    const int32_t *push_ptr;     /* HOST [n_bpatch+1] CSR over those patches into the push entries     */
    const int32_t *push_cell;    /* HOST [n_entries] cell index inside its patch, one entry per (cell, peer) */
    int32_t n_recv, n_send;      /* <= 16 each                                                         */
    int32_t recv_peer[16];       /* ranks this rank receives ghosts from (index into flags)            */
    uint64_t remote_flag[16];    /* DEVICE address of flags[this rank] on every rank this rank sends to */
    uint64_t flags;              /* DEVICE address of this rank's uint64 flags[world], zero-initialised, peer-writable */
} tb_halo_fused;
int tb_halo_fused_setup(tb_ctx *ctx, const tb_halo_fused *h);
/* tb_swe_stage over all patches + push: push_dst is a DEVICE array [n_entries] of peer addresses (uint64) of the
 * pushed records inside the peers' copies of THIS output buffer.  Every rank must issue the same sequence of fused
 * launches. */
int tb_swe_stage_fused(tb_ctx *ctx, double a0, double a1, double b_dt, const double *u_in, const double *u0,
                       double *u_out, const uint64_t *push_dst, void *stream);
/* The same for the tracer stage and the limiter (3 doubles per pushed cell): push_dst holds the peer addresses of the
 * pushed cells inside the peers' copies of c_out.  SWE, tracer and limiter launches share ONE epoch sequence: every
 * rank issues the same sequence of fused launches, and the boundary patches of launch k+1 wait for the peers' launch k
 * (the tracer stage thereby also sees the frozen SWE ghosts of the last SWE stage). */
int tb_tracer_stage_fused(tb_ctx *ctx, double a0, double a1, double b_dt, const double *c_in, const double *c0,
                          double *c_out, const double *swe_state, const uint64_t *push_dst, void *stream);
int tb_limiter_apply_to_fused(tb_ctx *ctx, const double *c_in, double *c_out, const uint64_t *push_dst, void *stream);
/* Stream-ordered wait until the ghost records of the last fused launch have arrived from every peer: required in
 * front of any other kernel that reads them (the tracer stage reads the frozen SWE state of its halo cells). */
int tb_halo_fused_wait(tb_ctx *ctx, void *stream);
/* Fused launches completed so far and whether a flag wait ever timed out (a peer stopped).  Synchronises. */
int tb_halo_fused_status(tb_ctx *ctx, int64_t *epoch, int32_t *error);

/* Restrict the next tb_swe_stage / tb_tracer_stage launches to patches
 * [first, first+count) (interior / partition-boundary split for overlap).
 * count < 0 resets to all patches. */
int tb_set_patch_range(tb_ctx *ctx, int64_t first, int64_t count);

/* Diagnostics: evaluates the kernels' fp64 helpers on DEVICE x[n]: out[0:n] = x^-1/2, out[n:2n] = sqrt(x),
 * out[2n:3n] = 1/x, out[3n:4n] = x^-1/3 (accuracy is asserted in tests/test_gpu_math.py). */
int tb_selftest_math(tb_ctx *ctx, const double *x, double *out, int64_t n, void *stream);

/* Same with an explicit DEVICE list of patch ids (n > 0), e.g. the patches that hold cells a peer needs first,
 * then the rest while the halo is in flight.  n == 0 resets to all patches.  SWE stage only. */
int tb_set_patch_list(tb_ctx *ctx, const int32_t *list, int64_t n);

/* launches issued by this library so far (bench.py's gpu_launches) */
int64_t tb_launch_count(const tb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* THETIS_B200_H */
