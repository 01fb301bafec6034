// Device-side helpers shared by the stage kernels (tb_kernels.cu, tb_tracer.cu): TMA bulk copies + mbarrier,
// cp.async, L2 prefetch and the branch-free fp64 reciprocal / root helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// 2-point Gauss-Legendre on [0,1]
#define TB_XI1 0.21132486540518711775
#define TB_XI2 0.78867513459481288225

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ fp64 math helpers
// MUFU / fp32 seed (>= 21 good bits) + ONE third-order (Halley-type) correction: with e = 1 - x*y0^k,
//   x^(-1/2) = y0 (1 + e/2 + 3e^2/8 + O(e^3)),  x^(-1) = y0 (1 + e + e^2 + O(e^3)),  x^(-1/3) = y0 (1 + e/3 + 2e^2/9 + O(e^3));
// the dropped e^3 term is < 2^-60, so results are good to ~1 ulp (tests/test_gpu_math.py), branch-free, all FMAs.
// The CUDA library sqrt / division / rcbrt carry slow-path subroutine calls that cost more fp64-pipe and issue
// slots than the whole facet flux.  Arguments here are depths, lengths and areas (normal, positive).
__device__ __forceinline__ double tb_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}
__device__ __forceinline__ double tb_sqrt(double x) {   // x > 0
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double g = x * y;
    const double e = fma(-g, y, 1.0);
    return fma(g * e, fma(0.375, e, 0.5), g);
}
__device__ __forceinline__ double tb_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ double tb_rcbrt(double x) {   // x^(-1/3), x > 0 within float range
    // seed 2^(-log2(x)/3) from the two MUFU approximations (relative error < 2^-20 for 1e-6 < x < 1e12; the library
    // rcbrtf costs ~19 instructions, 6 % of everything the config-5 stage kernel executes)
    float lg, y0;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float)x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(lg * (-1.0f / 3.0f)));
    const double y = (double)y0;
    const double e = fma(-(x * y), y * y, 1.0);
    return fma(y * e, fma(2.0 / 9.0, e, 1.0 / 3.0), y);
}

__device__ __forceinline__ void cp_async8(void *sdst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *gsrc) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}


// ------------------------------------------------------------------ cross-GPU flags (system scope)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}


// ------------------------------------------------------------------ fused halo exchange (TbHaloFused, tb_internal.h)
// Shared by the SWE stage, tracer stage and limiter kernels: every rank issues the same sequence of fused launches and
// one epoch counter orders them all.
#ifdef TB_HAVE_HALO_FUSED
// Partition-boundary CTA, before it reads ghost records: wait until every peer this rank receives from has published
// the ghosts written by ITS previous fused launch.  Returns the epoch of the running launch.
__device__ __forceinline__ unsigned long long tb_fused_wait(const TbHaloFused *hf) {
    const unsigned long long epoch = *hf->epoch;
    // fail fast: once a wait has timed out (a peer stopped, or the ranks did not issue the same sequence of fused
    // launches) no later launch spins again -- the host reads the flag (tb_halo_fused_status) and aborts the run
    if (*reinterpret_cast<volatile int *>(hf->error)) return epoch;
    const long long t0 = clock64();
    for (int q = 0; q < hf->n_recv; ++q) {
        const unsigned long long *f = hf->flags + hf->recv_peer[q];
        while (ld_acquire_sys(f) < epoch) {
            if (clock64() - t0 > 6000000000ll) {      // ~3 s: a peer died; do not hang the GPU
                *hf->error = 1;
                break;
            }
        }
    }
    return epoch;
}
// Partition-boundary CTA, after its results are staged in shared memory (O: [TB_P][REC] doubles): store the records
// the peers need straight into their ghost blocks (consecutive threads store consecutive doubles of a record) and let
// the last boundary CTA of the launch publish the new epoch to the receiving peers.  All threads must call it.
template <int REC>
__device__ __forceinline__ void tb_fused_push(const TbHaloFused *hf, const unsigned long long *push_dst, const double *O,
                                              int n_bpatch, unsigned long long epoch, int tid) {
    const int e0 = __ldg(hf->push_ptr + blockIdx.x), n = (__ldg(hf->push_ptr + blockIdx.x + 1) - e0) * REC;
    for (int i = tid; i < n; i += TB_P) {
        const int ent = i / REC, k = i - ent * REC;
        double *dst = reinterpret_cast<double *>(__ldg(push_dst + e0 + ent));
        dst[k] = O[__ldg(hf->push_cell + e0 + ent) * REC + k];
    }
    __threadfence_system();          // this thread's peer stores are ordered before the signal below
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();      // cumulative over the whole CTA's stores (observed through the barrier)
        const unsigned int done = atomicAdd(hf->done_count, 1u);
        if (done == (unsigned int)n_bpatch - 1u) {
            // every boundary CTA of this launch has pushed: publish the new epoch to the receiving peers
            *hf->done_count = 0u;
            __threadfence_system();
            for (int q = 0; q < hf->n_send; ++q) st_release_sys(hf->remote_flag[q], epoch + 1ull);
            *hf->epoch = epoch + 1ull;
        }
    }
}
#endif
