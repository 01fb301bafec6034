// Explicit 2-D tracer advection stage (tracer_eq_2d.py:147-193, 293-298) and
// VertexBasedP1DGLimiter (limiter.py:48-198) as sm_100a kernels.
// Same patch / TMA staging scheme as the SWE stage kernel (tb_kernels.cu).
#include "tb_internal.h"
#include "tb_device.cuh"

#define TB_BC_PRESENT 32   // marker has a (possibly empty) dict in bnd_conditions
#ifndef TB_T_HALO_SPEC
#define TB_T_HALO_SPEC 4      // halo elements (9 per halo cell) per thread fetched speculatively (covers NH <= 56)
#endif
#ifndef TB_T_MINB
#define TB_T_MINB 7           // resident CTAs per SM the tracer stage kernel is compiled for (72 registers)
#endif

// degree-3 cell rule for the non-polynomial integrand of ConservativeSourceTerm (H*source with wetting-drying)
__constant__ double ct_qlam[TB_MAX_QUAD][3];
__constant__ double ct_qw[TB_MAX_QUAD];
cudaError_t tb_set_quadrature_tracer(int n, const double *lam, const double *w) {
    cudaError_t e = cudaMemcpyToSymbol(ct_qlam, lam, sizeof(double) * 3 * n);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(ct_qw, w, sizeof(double) * n);
}

// TSPEC 1: plain non-conservative advection (no source, diffusion or Lax-Friedrichs) -- BASELINE config 4;
// TSPEC 0: every optional term behind a runtime flag.  Same staging as the SWE stage kernel: TMA bulk copies of the
// patch's SWE records, tracer records, u0 and static block; halo records gathered with cp.async by all threads.
// shared-memory layout: barrier | S | C | O | static block | halo ids
__host__ __device__ inline size_t tb_tracer_ids_offset(const TbPatchLayout &pl) {
    return 16 + (size_t)(TB_P + pl.NH) * 96 + (size_t)TB_P * 24 + (size_t)pl.stride;
}

template <int TSPEC>
__global__ void __launch_bounds__(TB_P, TB_T_MINB) tracer_stage_kernel(const __grid_constant__ TbTracerParams prm) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    double *S = reinterpret_cast<double *>(smem + 16);            // SWE records [(TB_P+NH)][9]
    double *C = S + (size_t)(TB_P + prm.pl.NH) * 9;               // tracer [(TB_P+NH)][3]
    double *O = C + (size_t)(TB_P + prm.pl.NH) * 3;               // [TB_P][3]  (+pad to 16 B)
    unsigned char *blk = reinterpret_cast<unsigned char *>(O + TB_P * 3);

    const int tid = threadIdx.x;
    const int patch = prm.patch_list ? __ldg(prm.patch_list + blockIdx.x) : prm.patch_first + (int)blockIdx.x;
    const long long cell0 = (long long)patch * TB_P;
    const int NV = prm.pl.NV;
    // distributed run, partition-boundary patch: its ghost records (tracer AND frozen SWE state) were written by the
    // peers' previous fused launch (tb_fused_wait, tb_device.cuh)
    const bool bpatch = prm.halo != nullptr && (int)blockIdx.x < prm.n_bpatch;
    const unsigned long long epoch = bpatch ? tb_fused_wait(prm.halo) : 0ull;

    // halo ids of this patch: one coalesced load per thread, staged in shared memory for the gather below (as in the
    // SWE stage kernel).  Element i of the halo is double (i % 9) of halo cell (i / 9): the 6 velocity values of its
    // SWE record, then its 3 tracer values (the neighbour's elevation is never used).
    const int *hid = prm.pl.halo_ids + (long long)patch * prm.pl.NH;
    int *ids_s = reinterpret_cast<int *>(smem + tb_tracer_ids_offset(prm.pl));
    const int myid = tid < prm.pl.NH ? __ldg(hid + tid) : 0;
    const int nh9 = __ldg(prm.pl.halo_cnt + patch) * 9;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        const uint32_t sb = (uint32_t)prm.pl.stride;
        const uint32_t rec9 = TB_P * 9 * sizeof(double), rec3 = TB_P * 3 * sizeof(double);
        mbar_expect_tx(bar, rec9 + rec3 + sb + (prm.c0 ? rec3 : 0u));
        bulk_g2s(S, prm.swe + cell0 * 9, rec9, bar);
        bulk_g2s(C, prm.c_in + cell0 * 3, rec3, bar);
        bulk_g2s(blk, prm.pl.sblk + (long long)patch * prm.pl.stride, sb, bar);
        if (prm.c0) bulk_g2s(O, prm.c0 + cell0 * 3, rec3, bar);
    } else if (tid == 32) {
        // warm L2 for the patch that runs on this SM slot one wave later
        const int pb = (int)blockIdx.x + 148 * TB_T_MINB;
        if (pb < (int)gridDim.x) {
            const long long pf = prm.patch_list ? __ldg(prm.patch_list + pb) : prm.patch_first + pb;
            bulk_prefetch_l2(prm.swe + pf * TB_P * 9, TB_P * 9 * sizeof(double));
            bulk_prefetch_l2(prm.c_in + pf * TB_P * 3, TB_P * 3 * sizeof(double));
            bulk_prefetch_l2(prm.pl.sblk + pf * prm.pl.stride, (uint32_t)prm.pl.stride);
            if (prm.c0) bulk_prefetch_l2(prm.c0 + pf * TB_P * 3, TB_P * 3 * sizeof(double));
        }
    }
    if (tid < prm.pl.NH) ids_s[tid] = myid;
    for (int h = TB_P + tid; h < prm.pl.NH; h += TB_P) ids_s[h] = __ldg(hid + h);      // very large halos only
    __syncthreads();          // ids staged; mbarrier initialised before anybody waits on it
    {
        auto halo_copy = [&](unsigned i) {
            const unsigned h = i / 9u, k = i - h * 9u;
            const long long gc = ids_s[h];
            if (k < 6) cp_async8(S + (TB_P + h) * 9 + k, prm.swe + gc * 9 + k);
            else cp_async8(C + (TB_P + h) * 3 + (k - 6), prm.c_in + gc * 3 + (k - 6));
        };
#pragma unroll
        for (int j = 0; j < TB_T_HALO_SPEC; ++j) {
            const int i = j * TB_P + tid;
            if (i < nh9) halo_copy((unsigned)i);
        }
        for (int i = TB_T_HALO_SPEC * TB_P + tid; i < nh9; i += TB_P) halo_copy((unsigned)i);   // large halos
        cp_async_wait_all();
    }
    __syncthreads();          // every thread's halo copies have landed
    mbar_wait(bar, 0);

    const bool active = (cell0 + tid) < prm.n_owned;
    double res[3] = {0, 0, 0};
    if (active) {
        const double *cols = reinterpret_cast<const double *>(blk);
        const uint16_t *cv = reinterpret_cast<const uint16_t *>(blk + prm.pl.off_cv) + tid * 3;
        const int *cn = reinterpret_cast<const int *>(blk + prm.pl.off_cn) + tid * 3;
        const double *my = S + tid * 9;
        const double corr = prm.corr;
        double ux[3], uy[3], c[3], x[3], y[3];
        // elevation and bathymetry are needed by rare branches only ('flux' boundary data, conservative source): they
        // are read from shared memory there instead of being kept in registers
        auto et = [&](int a) { return my[6 + a]; };
        auto b = [&](int a) { return cols[2 * NV + cv[a]]; };
        int v[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            ux[a] = corr * my[2 * a];
            uy[a] = corr * my[2 * a + 1];
            c[a] = C[tid * 3 + a];
            v[a] = cv[a];
            x[a] = cols[v[a]];
            y[a] = cols[NV + v[a]];
        }
        double Nx[3], Ny[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int p = (i + 1) % 3, q = (i + 2) % 3;
            Nx[i] = y[q] - y[p];
            Ny[i] = x[p] - x[q];
        }
        const double twoA = (x[1] - x[0]) * (y[2] - y[0]) - (y[1] - y[0]) * (x[2] - x[0]);
        double R[3] = {0, 0, 0};
        const double sc = c[0] + c[1] + c[2];
        const bool cons = TSPEC == 1 ? false : prm.conservative != 0;   // depth-integrated unknown (tracer_eq_2d.py:323-437)
        const bool has_diff = TSPEC == 1 ? false : prm.diff.mode != 0;
        const bool has_src = TSPEC == 1 ? false : prm.src.mode != 0;
        const bool lf_on = TSPEC == 1 ? false : prm.lf_on != 0;
        {
            // cell part (:159-160): + c div(u phi_a); conservative form (:356-357): + c u.grad(phi_a)
            double D = 0;
            if (!cons) {
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) D += -0.5 * (Nx[bb] * ux[bb] + Ny[bb] * uy[bb]);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double s = D * (c[a] + sc);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) s += -0.5 * (Nx[a] * ux[cc] + Ny[a] * uy[cc]) * (c[cc] + sc);
                R[a] += s * (1.0 / 12.0);
            }
        }
        if (has_src) {
            // SourceTerm (:293-298)
            double f[3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
                f[a] = prm.src.mode == 3 ? __ldg(prm.src.cell + (cell0 + tid) * 3 + a)       // P1DG source (tb_set_field_cell)
                       : prm.src.mode == 2 ? cols[(size_t)prm.src.col * NV + v[a]] : prm.src.v0;
            const double s = f[0] + f[1] + f[2];
            if (!cons) {
#pragma unroll
                for (int a = 0; a < 3; ++a) R[a] += 0.5 * twoA * (1.0 / 12.0) * (f[a] + s);
            } else {
                // ConservativeSourceTerm (:429-437): int H*source*phi_a by the cell rule
                double hl[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) hl[a] = prm.nonlin ? b(a) + et(a) : b(a);
                for (int qd = 0; qd < prm.nquad; ++qd) {
                    const double l0 = ct_qlam[qd][0], l1 = ct_qlam[qd][1], l2 = ct_qlam[qd][2];
                    double Hq = l0 * hl[0] + l1 * hl[1] + l2 * hl[2];
                    if (prm.nonlin && prm.wd_on) Hq = 0.5 * (Hq + sqrt(Hq * Hq + prm.wd_alpha2));
                    const double k = ct_qw[qd] * 0.5 * twoA * Hq * (l0 * f[0] + l1 * f[1] + l2 * f[2]);
                    R[0] += l0 * k; R[1] += l1 * k; R[2] += l2 * k;
                }
            }
        }
        // HorizontalDiffusionTerm (tracer_eq_2d.py:226-278), symmetric interior penalty
        double muv[3] = {0, 0, 0}, gcx = 0, gcy = 0;
        const double itA = tb_rcp(twoA);
        if (has_diff) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                muv[a] = prm.diff.mode == 2 ? cols[(size_t)prm.diff.col * NV + v[a]] : prm.diff.v0;
                gcx -= itA * c[a] * Nx[a];        // grad(phi_a) = -N_a/2A
                gcy -= itA * c[a] * Ny[a];
            }
            // cell term (:238): -int mu grad(phi_a).grad(c) = +1/2 mean(mu) N_a.grad(c)
            const double hm = (1.0 / 6.0) * (muv[0] + muv[1] + muv[2]);
#pragma unroll
            for (int a = 0; a < 3; ++a) R[a] += hm * (Nx[a] * gcx + Ny[a] * gcy);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int p = (i + 1) % 3, q = (i + 2) % 3;
            const double nxs = Nx[i], nys = Ny[i];
            const int code = cn[i];
            double Fp = 0, Fq = 0;
            if (code >= 0) {
                const int lf = code & 3;
                const int ni = code >> 2;
                const double *nr = S + ni * 9;
                const double *nc = C + ni * 3;
                const int np_ = (lf + 2) % 3, nq_ = (lf + 1) % 3;
                const double uNxp = corr * nr[2 * np_], uNyp = corr * nr[2 * np_ + 1], cNp = nc[np_];
                const double uNxq = corr * nr[2 * nq_], uNyq = corr * nr[2 * nq_ + 1], cNq = nc[nq_];
                double sigl = 0, T = 0, Dc = 0;
                if (has_diff) {
                    // neighbour geometry -> its grad(c) and area (:241-258)
                    const uint16_t *cvn = ni < TB_P
                        ? reinterpret_cast<const uint16_t *>(blk + prm.pl.off_cv) + ni * 3
                        : reinterpret_cast<const uint16_t *>(blk + prm.pl.off_hcv) + (ni - TB_P) * 3;
                    double xn[3], yn[3];
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) {
                        xn[bb] = cols[cvn[bb]];
                        yn[bb] = cols[NV + cvn[bb]];
                    }
                    const double twoAN = (xn[1] - xn[0]) * (yn[2] - yn[0]) - (yn[1] - yn[0]) * (xn[2] - xn[0]);
                    double hx = 0, hy = 0;
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) {
                        hx += nc[bb] * (yn[(bb + 2) % 3] - yn[(bb + 1) % 3]);
                        hy += nc[bb] * (xn[(bb + 1) % 3] - xn[(bb + 2) % 3]);
                    }
                    hx /= -twoAN;
                    hy /= -twoAN;
                    T = (gcx + hx) * nxs + (gcy + hy) * nys;
                    sigl = 6.0 * prm.sipg * (nxs * nxs + nys * nys) / fmin(twoA, twoAN);   // sigma_max * len
                }
#pragma unroll
                for (int gp = 0; gp < 2; ++gp) {
                    const double wq_ = gp ? TB_XI2 : TB_XI1, wp_ = 1.0 - wq_;
                    const double uKx = wp_ * ux[p] + wq_ * ux[q], uKy = wp_ * uy[p] + wq_ * uy[q];
                    const double uNx = wp_ * uNxp + wq_ * uNxq, uNy = wp_ * uNyp + wq_ * uNyq;
                    const double cK = wp_ * c[p] + wq_ * c[q], cN = wp_ * cNp + wq_ * cNq;
                    const double unav = 0.5 * ((uKx + uNx) * nxs + (uKy + uNy) * nys);   // avg(u).n * len (own outward n)
                    // upwind value (:164-168); sign(0) = 0 gives the mean
                    double f;
                    if (!cons) {
                        const double cup = unav > 0.0 ? cK : (unav < 0.0 ? cN : 0.5 * (cK + cN));
                        f = cup * (uKx * nxs + uKy * nys);                                 // own velocity trace (:170-171)
                    } else {
                        // flux_up = c u of the upwind side (:364-370)
                        const double fK = cK * (uKx * nxs + uKy * nys), fN = cN * (uNx * nxs + uNy * nys);
                        f = unav > 0.0 ? fK : (unav < 0.0 ? fN : 0.5 * (fK + fN));
                    }
                    if (lf_on) f += 0.5 * fabs(unav) * prm.lf_sigma * (cK - cN);         // (:173-175, 372-380)
                    if (has_diff) {
                        const double mug = wp_ * muv[p] + wq_ * muv[q];                    // mu is continuous (P1)
                        f += mug * (sigl * (cK - cN) - 0.5 * T);
                        Dc += 0.5 * mug * (cK - cN);
                    }
                    Fp += wp_ * f;
                    Fq += wq_ * f;
                }
                if (has_diff) {
                    // -inner(avg(mu grad(phi)), jump(c, n)) (:253-254): all three nodes
#pragma unroll
                    for (int a = 0; a < 3; ++a) R[a] -= 0.5 * itA * (Nx[a] * nxs + Ny[a] * nys) * Dc;
                }
            } else {
                const int gb = -(code + 1);
                const int slot = __ldg(prm.bc.bf_slot + gb);
                const TbBcSlot &bs = prm.bc.slots[slot];
                const int op = bs.opcode;
                const int row = bs.arr_mask ? __ldg(prm.bc.bf_row + gb) : 0;
                const double len2 = nxs * nxs + nys * nys;
                const double il = tb_rsqrt(len2);
#pragma unroll
                for (int gp = 0; gp < 2; ++gp) {
                    const double wq_ = gp ? TB_XI2 : TB_XI1, wp_ = 1.0 - wq_;
                    const double uKx = wp_ * ux[p] + wq_ * ux[q], uKy = wp_ * uy[p] + wq_ * uy[q];
                    const double cK = wp_ * c[p] + wq_ * c[q];
                    double f;
                    if (!(op & TB_BC_PRESENT)) {
                        f = cK * (uKx * nxs + uKy * nys);                                 // closed (:189-191)
                    } else {
                        // TracerTerm.get_bnd_functions (:78-115)
                        double cext = cK, uex = uKx, uey = uKy;
                        if (op & TB_BC_VALUE) {
                            cext = bs.value;
                            if (bs.arr_mask & TB_BC_VALUE)
                                cext = wp_ * __ldg(prm.bc.ext_value + 2 * row) + wq_ * __ldg(prm.bc.ext_value + 2 * row + 1);
                        }
                        if (op & TB_BC_UV) {
                            double uvx = bs.uvx, uvy = bs.uvy;
                            if (bs.arr_mask & TB_BC_UV) {
                                uvx = wp_ * __ldg(prm.bc.ext_uv + 4 * row) + wq_ * __ldg(prm.bc.ext_uv + 4 * row + 2);
                                uvy = wp_ * __ldg(prm.bc.ext_uv + 4 * row + 1) + wq_ * __ldg(prm.bc.ext_uv + 4 * row + 3);
                            }
                            uex = corr * uvx;
                            uey = corr * uvy;
                        } else if (op & TB_BC_FLUX) {
                            double flux = bs.flux;
                            if (bs.arr_mask & TB_BC_FLUX)
                                flux = wp_ * __ldg(prm.bc.ext_flux + 2 * row) + wq_ * __ldg(prm.bc.ext_flux + 2 * row + 1);
                            double eext = wp_ * et(p) + wq_ * et(q);
                            if (op & TB_BC_ELEV) {
                                eext = bs.elev;
                                if (bs.arr_mask & TB_BC_ELEV)
                                    eext = wp_ * __ldg(prm.bc.ext_elev + 2 * row) + wq_ * __ldg(prm.bc.ext_elev + 2 * row + 1);
                            }
                            const double bg = wp_ * b(p) + wq_ * b(q);
                            double hext = bg;
                            if (prm.nonlin) {
                                hext = bg + eext;
                                if (prm.wd_on) hext = 0.5 * (hext + sqrt(hext * hext + prm.wd_alpha2));
                            }
                            const double s = corr * flux / (hext * bs.bnd_len);
                            uex = s * nxs * il;
                            uey = s * nys * il;
                        } else if (op & TB_BC_UN) {
                            double un = bs.un;
                            if (bs.arr_mask & TB_BC_UN)
                                un = wp_ * __ldg(prm.bc.ext_un + 2 * row) + wq_ * __ldg(prm.bc.ext_un + 2 * row + 1);
                            uex = un * nxs * il;
                            uey = un * nys * il;
                        }
                        const double unav = 0.5 * ((uKx + uex) * nxs + (uKy + uey) * nys);
                        if (!cons) {
                            const double cup = unav > 0.0 ? cK : (unav < 0.0 ? cext : 0.5 * (cK + cext));
                            f = cup * unav;                                                // (:181-188)
                        } else {
                            // flux_up = c_in*uv*s + c_ext*uv_ext*(1-s)  (:391-394)
                            const double fK = cK * (uKx * nxs + uKy * nys), fE = cext * (uex * nxs + uey * nys);
                            f = unav > 0.0 ? fK : (unav < 0.0 ? fE : 0.5 * (fK + fE));
                        }
                        if (has_diff) {
                            if (op & TB_BC_DIFF_FLUX) {
                                f -= bs.diff_flux * len2 * il;                              // (:264-265)
                            } else {
                                // -test*dot(mu grad(c_up), n) (:267-272): grad(c_ext) = 0 for a Constant 'value'
                                const double sg = (op & TB_BC_VALUE) ? (unav > 0.0 ? 1.0 : (unav < 0.0 ? 0.0 : 0.5)) : 1.0;
                                f -= sg * (wp_ * muv[p] + wq_ * muv[q]) * (gcx * nxs + gcy * nys);
                            }
                        }
                    }
                    Fp += wp_ * f;
                    Fq += wq_ * f;
                }
            }
            R[p] -= 0.5 * Fp;
            R[q] -= 0.5 * Fq;
        }
        const double mi = 6.0 * itA * prm.bdt;
        const double sR = R[0] + R[1] + R[2];
#pragma unroll
        for (int a = 0; a < 3; ++a) res[a] = prm.a1 * c[a] + mi * (4.0 * R[a] - sR);
        if (prm.c0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) res[a] += prm.a0 * O[tid * 3 + a];
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) O[tid * 3 + a] = res[a];
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(prm.c_out + cell0 * 3, O, TB_P * 3 * sizeof(double));
        bulk_commit_wait_read();
    }
    // fused halo push of the new tracer values (3 doubles per cell) into the peers' ghost blocks
    if (bpatch) tb_fused_push<3>(prm.halo, prm.push_dst, O, prm.n_bpatch, epoch, tid);
}

size_t tb_tracer_smem_bytes(const TbPatchLayout &pl) {
    return tb_tracer_ids_offset(pl) + (((size_t)pl.NH * sizeof(int) + 15) & ~(size_t)15);
}

cudaError_t tb_tracer_kernels_init() {
    cudaError_t e = cudaFuncSetAttribute(tracer_stage_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tracer_stage_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

cudaError_t tb_launch_tracer_stage(const TbTracerParams &p, int n_patches, size_t smem, cudaStream_t s) {
    if (n_patches <= 0) return cudaSuccess;
    const bool plain = !p.conservative && !p.diff.mode && !p.src.mode && !p.lf_on && !p.force_generic;
    if (plain) tracer_stage_kernel<1><<<n_patches, TB_P, smem, s>>>(p);
    else tracer_stage_kernel<0><<<n_patches, TB_P, smem, s>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ limiter
// VertexBasedP1DGLimiter.apply (limiter.py:182-198 over firedrake's VertexBasedLimiter) as ONE patch-staged kernel.
// One CTA = one patch of TB_P cells (the patches of the stage kernels).  Per patch a static table (built once,
// limiter_setup in tb_api.cu) lists the cells outside the patch that share a vertex with it ("vertex halo"), the
// patch-local vertex of each own-cell node, and -- per patch vertex -- the cells around it (CSR over shared-memory
// slots: own cells 0..TB_P-1, halo cells TB_P..) plus the exterior facets touching it.
//   (i)   every thread loads its cell (24 B, read once) into shared memory together with its P0 projection = mean of
//         the nodal values (limiter.py:90-97); the vertex-halo cells are gathered by the same threads (mostly L2
//         hits: they belong to patches running at the same time);
//   (ii)  one thread per patch vertex: min / max over the means of ALL cells around it, plus the mean of the two nodal
//         values of every exterior facet touching it (limiter.py:109-145) -- a deterministic gather from shared
//         memory, no atomics (a first version with shared-memory atomic min / max was bound by the atomic unit:
//         ~2 400 atomics per patch);
//   (iii) per-cell clamp (VertexBasedLimiter._limit_kernel) and ONE write of the limited cell (24 B).
// Out of place (c_in -> c_out): neighbouring patches read this patch's ORIGINAL values as their vertex halo.
// Replaces the two global passes (vertex CSR gather + per-cell clamp) that moved 94 B per triangle for 64 needed.
#ifndef TB_LIM_MINB
#define TB_LIM_MINB 10      // resident CTAs per SM the limiter kernel is compiled for (the kernel is latency bound)
#endif
#ifndef TB_LIM_PREFETCH
#define TB_LIM_PREFETCH (148 * TB_LIM_MINB)      // patches ahead to warm in L2: one wave
#endif
#ifndef TB_LIM_SPEC
#define TB_LIM_SPEC 2       // vertex-halo cells per thread gathered speculatively (covers NHV <= 2 TB_P)
#endif

__global__ void __launch_bounds__(TB_P, TB_LIM_MINB) limiter_patch_kernel(TbLimiterData d, const double *__restrict__ c_in,
                                                                          double *__restrict__ c_out) {
    extern __shared__ __align__(16) unsigned char lsm[];
    double *qs = reinterpret_cast<double *>(lsm);                   // [(TB_P + NHV)][3] nodal values, own + halo
    double *ms = qs + (size_t)(TB_P + d.NHV) * 3;                   // [(TB_P + NHV)] cell means
    double *qmin_s = ms + (TB_P + d.NHV);                           // [NVT]
    double *qmax_s = qmin_s + d.NVT;
    const int tid = threadIdx.x;
    const long long patch = d.patch_list ? __ldg(d.patch_list + blockIdx.x) : (long long)blockIdx.x;
    const long long cell = patch * TB_P + tid;
    const bool active = cell < d.n_owned;
    // distributed run, partition-boundary patch: its vertex-halo ghosts were written by the peers' previous fused launch
    const bool bpatch = d.halo != nullptr && (int)blockIdx.x < d.n_bpatch;
    const unsigned long long epoch = bpatch ? tb_fused_wait(d.halo) : 0ull;
    const unsigned char *blk = d.tab + patch * d.stride;
    const int *hids = reinterpret_cast<const int *>(blk);
    const unsigned short *ctv = reinterpret_cast<const unsigned short *>(blk + d.off_ctv);
    const unsigned short *vptr = reinterpret_cast<const unsigned short *>(blk + d.off_vptr);
    const unsigned short *vidx = reinterpret_cast<const unsigned short *>(blk + d.off_vidx);
    const int2 cnt = __ldg(reinterpret_cast<const int2 *>(d.counts) + patch);      // (vertex-halo cells, vertices)
    const int nhv = cnt.x, nvt = cnt.y;

    // warm L2 for the patch that will run on this SM slot one wave later (the kernel is bound by the latency of its
    // dependent loads -- table -> ids -> gathered values -- not by bandwidth)
    if (tid < 2) {
        const long long pb = (long long)blockIdx.x + (long long)TB_LIM_PREFETCH;
        if (pb < (long long)gridDim.x) {
            const long long pf = d.patch_list ? __ldg(d.patch_list + pb) : pb;
            if (tid == 0) bulk_prefetch_l2(c_in + pf * TB_P * 3, TB_P * 3 * sizeof(double));
            else bulk_prefetch_l2(d.tab + pf * d.stride, (uint32_t)d.stride);
        }
    }
    // every independent global load first (one memory latency instead of a chain): the ids of the halo cells this
    // thread will gather (rows are padded to NHV valid entries), its own cell and its table entries
    int hid[TB_LIM_SPEC];
#pragma unroll
    for (int j = 0; j < TB_LIM_SPEC; ++j) {
        const int h = j * TB_P + tid;
        hid[j] = h < d.NHV ? __ldg(hids + h) : 0;
    }
    double q[3] = {0, 0, 0};
    int lv[3] = {0, 0, 0};
    if (active) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            q[a] = __ldg(c_in + cell * 3 + a);
            lv[a] = ctv[tid * 3 + a];
        }
    }
    const double qavg = (q[0] + q[1] + q[2]) / 3.0;        // P0 projection = mean of the nodal values (limiter.py:90-97)
    qs[tid * 3] = q[0]; qs[tid * 3 + 1] = q[1]; qs[tid * 3 + 2] = q[2];
    ms[tid] = qavg;
    // ... then the gathers that depend on the ids
#pragma unroll
    for (int j = 0; j < TB_LIM_SPEC; ++j) {
        const int h = j * TB_P + tid;
        if (h < nhv) {
            const double *r = c_in + (long long)hid[j] * 3;
            const double h0 = __ldg(r), h1 = __ldg(r + 1), h2 = __ldg(r + 2);
            qs[(TB_P + h) * 3] = h0; qs[(TB_P + h) * 3 + 1] = h1; qs[(TB_P + h) * 3 + 2] = h2;
            ms[TB_P + h] = (h0 + h1 + h2) / 3.0;
        }
    }
    for (int h = TB_LIM_SPEC * TB_P + tid; h < nhv; h += TB_P) {        // very large vertex halos only
        const double *r = c_in + (long long)__ldg(hids + h) * 3;
        const double h0 = __ldg(r), h1 = __ldg(r + 1), h2 = __ldg(r + 2);
        qs[(TB_P + h) * 3] = h0; qs[(TB_P + h) * 3 + 1] = h1; qs[(TB_P + h) * 3 + 2] = h2;
        ms[TB_P + h] = (h0 + h1 + h2) / 3.0;
    }
    __syncthreads();
    // ---- vertex bounds: one thread per patch vertex, gather over the cells / exterior facets around it
    for (int v = tid; v < nvt; v += TB_P) {
        double qmax = -1.0e10, qmin = 1.0e10;   // firedrake VertexBasedLimiter.compute_bounds initial values
        const int k1 = vptr[v + 1];
        for (int k = vptr[v]; k < k1; ++k) {
            const int ent = vidx[k];
            double val;
            if (ent & 0x8000) {
                // exterior facet f of the cell in slot s touches this vertex: mean of its two nodal values (:123-137)
                const int sl = (ent & 0x7fff) >> 2, f = ent & 3;
                val = (qs[sl * 3 + (f + 1) % 3] + qs[sl * 3 + (f + 2) % 3]) / 2;
            } else {
                val = ms[ent];
            }
            qmax = fmax(qmax, val);
            qmin = fmin(qmin, val);
        }
        qmin_s[v] = qmin;
        qmax_s[v] = qmax;
    }
    __syncthreads();
    // ---- per-cell clamp (VertexBasedLimiter._limit_kernel) and ONE write of the limited cell
    double alpha = 1.0;
    if (active) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (q[i] > qavg)
                alpha = fmin(alpha, fmin(1.0, (qmax_s[lv[i]] - qavg) / (q[i] - qavg)));
            else if (q[i] < qavg)
                alpha = fmin(alpha, fmin(1.0, (qavg - qmin_s[lv[i]]) / (qavg - q[i])));
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            q[i] = qavg + alpha * (q[i] - qavg);
            c_out[cell * 3 + i] = q[i];
        }
    }
    if (bpatch) {
        // fused halo push of the limited values: staged over this patch's own slots of qs (every thread is past
        // the bounds phase: the barrier above)
#pragma unroll
        for (int i = 0; i < 3; ++i) qs[tid * 3 + i] = q[i];
        __syncthreads();
        tb_fused_push<3>(d.halo, d.push_dst, qs, d.n_bpatch, epoch, tid);
    }
}
cudaError_t tb_launch_limiter(const TbLimiterData &d, const double *c_in, double *c_out, cudaStream_t s) {
    const long long np = (d.n_owned + TB_P - 1) / TB_P;
    const size_t smem = ((size_t)(TB_P + d.NHV) * 4 + (size_t)d.NVT * 2) * sizeof(double);
    if (np > 0) limiter_patch_kernel<<<(unsigned)np, TB_P, smem, s>>>(d, c_in, c_out);
    return cudaGetLastError();
}
