// Rejected experiment (round 2): persistent k_final whose first-level records arrive by cp.async.bulk + mbarrier.
// Measured at C4 (B200): k_final 0.440 ms (256-thread CTAs) / 0.403 ms (64-thread CTAs) -> k_final_pf 0.455 (256) / 0.438 (128) / 0.430 (64) / 0.432 (32):
// profiles/experiments/r02v_bench_pf.json, r02w_bench_pf{128,64,32}.json.  Parity green (58 GPU tests).  Kept for reference; it
// plugs into interact_kernels.cuh after k_final (final_cell() is the shared body) with the launch
//   k_final_pf<CTAS><<<min(ceil(nC / NT), n_sm * CTAS), NT, 0, st>>>(I);

// ------------------------------------------------------------------------------------------------
// k_final_pf: the same pass as a persistent kernel whose first-level records — caller labels, slot counts, slot 0: the three
// streams every cell reads, 9 bytes per cell — arrive by bulk asynchronous copy (cp.async.bulk + mbarrier, the TMA engine's 1-D
// path), one chunk of FINAL_PF_NT positions ahead.  A third of k_final's stall samples are warps waiting for exactly these loads at
// their very first instruction; here they are in shared memory when the chunk starts, and the copies hold no registers (a
// register-held software prefetch of the same loads cost occupancy and lost, profiles/experiments/).
// Needs c_begin % 16 == 0 and FINAL_PF_NT % 16 == 0 (16-byte source alignment of the n_item stream); the last, partial chunk reads its records directly.
// ------------------------------------------------------------------------------------------------
#ifndef FINAL_PF_NT
#define FINAL_PF_NT 256
#endif
struct __align__(128) FinalStage {
    int orig[FINAL_PF_NT];
    int slot0[FINAL_PF_NT];
    unsigned char n_item[FINAL_PF_NT];
};
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok = 0;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int MINB>
__global__ void __launch_bounds__(FINAL_PF_NT, MINB) k_final_pf(InteractParams P) {
    __shared__ FinalStage st[2];
    __shared__ __align__(8) unsigned long long bar[2];
    constexpr int NT = FINAL_PF_NT;
    const int n_chunks = (P.c_end - P.c_begin + NT - 1) / NT;
    int k = blockIdx.x;
    if (k >= n_chunks) return;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // thread 0: request the records of chunk kk into stage b (a partial last chunk is read directly: the barrier just completes)
    auto request = [&](int kk, int b) {
        const int start = P.c_begin + kk * NT;
        if (P.c_end - start >= NT) {
            mbar_expect_tx(&bar[b], NT * 4 + NT * 4 + NT);
            bulk_g2s(st[b].orig, P.m.orig + start, NT * 4, &bar[b]);
            bulk_g2s(st[b].slot0, P.slots + start, NT * 4, &bar[b]);
            bulk_g2s(st[b].n_item, P.n_item + start, NT, &bar[b]);
        } else mbar_arrive(&bar[b]);
    };
    if (tid == 0) request(k, 0);
    int b = 0;
    unsigned parity = 0;      // bit b = the phase stage b completes next
    for (; k < n_chunks; k += gridDim.x) {
        if (tid == 0 && k + (int)gridDim.x < n_chunks) request(k + gridDim.x, b ^ 1);
        mbar_wait(&bar[b], (parity >> b) & 1u);
        parity ^= 1u << b;
        const int start = P.c_begin + k * NT, c = start + tid;
        const bool live = c < P.c_end;
        int oc = 0, n = 0, e0 = 0;
        if (P.c_end - start >= NT) { oc = st[b].orig[tid]; n = st[b].n_item[tid]; e0 = st[b].slot0[tid]; }
        else if (live) { oc = __ldg(P.m.orig + c); n = P.n_item[c]; e0 = P.slots[c]; }
        final_cell(P, c, live, oc, n, e0);
        __syncthreads();          // every thread has read stage b: it may be refilled (requested at the top of the next iteration)
        b ^= 1;
    }
}

