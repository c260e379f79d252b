// Rejected experiment (round 2): k_final with the touched cells compacted per warp (32 * FINAL_C_H positions per warp).
// Measured at C4 (B200): k_final (2-warp CTAs) 0.4047 ms; k_final_c 0.4506 (H = 1) / 0.4041 (H = 2) / 0.4153 (H = 4) / 0.4763 (H = 8):
// profiles/experiments/r02z_bench_fc*.json.  Parity green (58 GPU tests).  Halving the warps that run the long path changes
// nothing: the pass is not bound by that path but by its 1.69 GB of scattered 64-byte-run traffic (4.2 TB/s = 0.64 of the
// measured peak; k_apply_forcing, the same gather / scatter pattern with no dependent chain, reaches 0.70).
// Plugs into interact_kernels.cuh after k_final (final_cell() is the shared body); launch
//   k_final_c<32><<<ceil(nC / (32 * FINAL_C_H)), 32, 0, st>>>(I);

// ------------------------------------------------------------------------------------------------
// k_final_c: the same pass with the touched cells COMPACTED.  One warp owns 32 * FINAL_C_H consecutive positions: it reads their first-level
// records (FINAL_C_H per lane), writes the zeros of the untouched cells at once, and packs the touched ones — about half of them at C4,
// clustered — into a shared-memory list that the warp then works off 32 at a time.  The long path (cell geometry, U, queue
// results, solid records, pair terms, butterfly) therefore runs once per ~32 TOUCHED cells with every lane busy, instead of once
// per 32 positions with half the lanes idle: the pass is bound by resident warps x latency, and this halves the long-lived ones.
// ------------------------------------------------------------------------------------------------
#ifndef FINAL_C_H
#define FINAL_C_H 2
#endif
template <int MINB>
__global__ void __launch_bounds__(32, MINB) k_final_c(InteractParams P) {
    constexpr int H = FINAL_C_H, NP = 32 * H;     // positions per warp
    __shared__ int l_c[NP], l_oc[NP], l_e0[NP];
    __shared__ unsigned char l_n[NP];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int c0 = P.c_begin + blockIdx.x * NP;
    int c[H], oc[H], n[H], e0[H];
    bool live[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        c[h] = c0 + 32 * h + lane;
        live[h] = c[h] < P.c_end;
        oc[h] = live[h] ? __ldg(P.m.orig + c[h]) : 0;
        n[h] = live[h] ? (int)P.n_item[c[h]] : 0;
        e0[h] = live[h] ? P.slots[c[h]] : 0;
    }
    int cnt = 0;
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const bool touched = n[h] > 0;
        const unsigned mask = __ballot_sync(FULL, touched);
        if (touched) {
            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            l_c[pos] = c[h]; l_oc[pos] = oc[h]; l_e0[pos] = e0[h]; l_n[pos] = (unsigned char)n[h];
        } else if (live[h]) store_cell(P, oc[h], 0.0, D3{0.0, 0.0, 0.0}, 0.0, 0.0);
        cnt += __popc(mask);
    }
    __syncwarp();
    for (int base = 0; base < cnt; base += 32) {
        const int i = base + lane;
        const bool have = i < cnt;
        final_cell(P, have ? l_c[i] : 0, have, have ? l_oc[i] : 0, have ? (int)l_n[i] : 0, have ? l_e0[i] : 0);
    }
}

