// stream_search.cuh -- the latency regime (query batches of <= 64, BASELINE config 5) in ONE launch.
//
// `coarse_stream_kernel` (coarse_kernel.cuh) scores one corpus slab per launch, and the search
// alternates it with `refresh_threshold_kernel`: 3-4 slab launches + 2-3 refresh launches per
// batch.  On an 8-GPU shard the bytes of a batch take 0.55 ms and that launch structure another
// 0.1 ms: every slab launch pays its own prologue (TMEM allocation, barrier init, 96 KB of
// queries into shared memory, an empty TMA pipeline) and tail, every refresh a launch of 64 CTAs
// that the other 84 SMs sit out.
//
// Here the whole slab schedule runs inside one persistent kernel (one CTA per SM, all co-resident):
//   warp 0 / lane 0 : TMA producer -- queries once, then the corpus tiles of slab 0, 1, 2 ... with
//                     no pause at a slab boundary (the corpus does not depend on a threshold)
//   warp 1 / lane 0 : MMA issuer, likewise
//   warp 2          : TMEM allocator
//   warps 4..7      : filter epilogue (thread = corpus row) AND the threshold refresh between
//                     slabs: when the CTA has filtered its tiles of slab s it arrives on a
//                     grid-wide counter; CTA q (q < nq) waits for all arrivals, refreshes query
//                     q's list with its 128 epilogue threads (the same `refresh_list` the refresh
//                     kernel runs, on a named barrier) and bumps a second counter; every CTA's
//                     epilogue waits until that one has counted nq queries (ONE polling lane per
//                     CTA on one word: 4,700 lanes polling 64 per-query words cost 45 us per
//                     batch of 64), reloads the thresholds and goes on with slab s+1 -- whose first tiles are by then already in shared
//                     memory and TMEM, because the producer and the MMA issuer never stopped.
// Replaces, for nq <= 64, the per-slab launches of the same arithmetic: results are
// bit-identical (same scores, same thresholds, same lists up to the order of the appends).
//
// Deadlock safety: the grid never exceeds the SM count and a CTA takes a whole SM, so all CTAs
// are resident once the stream's previous kernel has drained.  Should something else keep CTAs
// from being scheduled (a shared GPU), every spin is bounded by wall time: the first waiter to
// time out raises `abort`, everybody stops waiting, and the kernel flags ALL queries of the batch
// as overflowed -- the host then answers them on the exact path.  Slow, never wrong, never hung.
#pragma once
#include "coarse_kernel.cuh"
#include "select_kernels.cuh"

namespace b2ip {

constexpr int STREAM_MAX_SLABS = 8;
constexpr int STREAM_REFRESH_KEYS = 4096;     // list staged in shared memory by the in-kernel refresh (32 KiB: the
                                              // stage count is not what limits the stream -- 5 to 12 stages measure the same)
constexpr int STREAM_REFRESH_BYTES = STREAM_REFRESH_KEYS * 8 + BOUND_BINS * 4 /*hist*/ + 64 /*prefix, krem, warp counts*/;
constexpr int STREAM_SYNC_WORDS = 2 + 64;     // [0] CTAs arrived at a slab end, [1] abort, [2] queries refreshed;
                                              // (the rest is spare; all zeroed by prep_queries_kernel)
using EpilogueGroup = WarpRangeGroup<128, 128, 1>;

struct StreamSearchParams {
    int n_slabs;
    long long slab_row[STREAM_MAX_SLABS + 1];   // slab s = local rows [slab_row[s], slab_row[s + 1])
    int dense_first;                  // slab 0 is stored densely (CoarseParams::dense)
    int k;
    float* thr;                       // [nq] thresholds (CoarseParams::thr, writable)
    int* kept;                        // [nq]
    const float* eps2;                // [nq]
    int* flags;                       // [nq]
    long long* gstats;
    unsigned int* sync;               // [STREAM_SYNC_WORDS], zeroed by prep_queries_kernel
    long long timeout_ns;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Waits until *word >= target (acquire, gpu scope).  false = the search was aborted (by this
// thread on a timeout, or by someone else).
__device__ __forceinline__ bool stream_wait(const unsigned int* word, unsigned int target,
                                            unsigned int* abort_word, long long timeout_ns) {
    if (ld_acquire_gpu(word) >= target) return true;
    const unsigned long long t0 = global_timer_ns();
    for (int spins = 0;; spins++) {
        __nanosleep(40);
        if (ld_acquire_gpu(word) >= target) return true;
        if ((spins & 15) == 15) {
            if (*reinterpret_cast<volatile unsigned int*>(abort_word) != 0u) return false;
            if (static_cast<long long>(global_timer_ns() - t0) > timeout_ns) {
                atomicExch(abort_word, 1u);
                return false;
            }
        }
    }
}

template <int NQ>
__global__ void __launch_bounds__(COARSE_THREADS, 1)
coarse_stream_search_kernel(const __grid_constant__ CUtensorMap tmap_q,
                            const __grid_constant__ CUtensorMap tmap_x, const CoarseParams p,
                            const __grid_constant__ StreamSearchParams sp, const int stages) {
    static_assert(NQ == 32 || NQ == 64, "query batch is padded to 32 or 64 columns");
    constexpr uint32_t kTmemCols = STREAM_ACC * NQ;
    constexpr int Q_KB_BYTES = NQ * KBLOCK_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smem_x = smem;                                           // [stages][128 x 128 B]
    uint8_t* smem_q = smem + stages * STREAM_X_STAGE_BYTES;           // [num_k_blocks][NQ x 128 B]
    uint8_t* smem_r = smem_q + p.num_k_blocks * Q_KB_BYTES;           // in-kernel refresh
    unsigned long long* r_keys = reinterpret_cast<unsigned long long*>(smem_r);
    unsigned int* r_hist = reinterpret_cast<unsigned int*>(smem_r + STREAM_REFRESH_KEYS * 8);
    unsigned long long* r_prefix = reinterpret_cast<unsigned long long*>(r_hist + BOUND_BINS);
    int* r_krem = reinterpret_cast<int*>(r_prefix + 1);
    int* r_warp = r_krem + 1;                                         // [8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_r + STREAM_REFRESH_BYTES);
    uint64_t* full_bar = bars;                                        // [STREAM_MAX_STAGES]
    uint64_t* empty_bar = bars + STREAM_MAX_STAGES;                   // [STREAM_MAX_STAGES]
    uint64_t* tfull_bar = bars + 2 * STREAM_MAX_STAGES;               // [STREAM_ACC]
    uint64_t* tempty_bar = bars + 2 * STREAM_MAX_STAGES + STREAM_ACC; // [STREAM_ACC]
    uint64_t* qfull_bar = bars + 2 * STREAM_MAX_STAGES + 2 * STREAM_ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qfull_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_q);
        ptx::prefetch_tmap(&tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STREAM_MAX_STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < STREAM_ACC; s++) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], 128);
        }
        ptx::mbar_init(qfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<kTmemCols>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer: all slabs back to back =====================
            ptx::mbar_expect_tx(qfull_bar, static_cast<uint32_t>(p.num_k_blocks * Q_KB_BYTES));
            for (int kb = 0; kb < p.num_k_blocks; kb++)
                ptx::tma_load_2d(smem_q + kb * Q_KB_BYTES, &tmap_q, qfull_bar, kb * KBLOCK_ELEMS, 0,
                                 p.hint_q);
            int stage = 0;
            uint32_t phase = 0;
            for (int s = 0; s < sp.n_slabs; s++) {
                const long long r0 = sp.slab_row[s];
                const int tiles = static_cast<int>((sp.slab_row[s + 1] - r0 + STREAM_TILE_X - 1) / STREAM_TILE_X);
                for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                    const long long x_row = r0 + static_cast<long long>(t) * STREAM_TILE_X;
                    for (int kb = 0; kb < p.num_k_blocks; kb++) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        ptx::mbar_expect_tx(&full_bar[stage], STREAM_X_STAGE_BYTES);
                        ptx::tma_load_2d(smem_x + stage * STREAM_X_STAGE_BYTES, &tmap_x, &full_bar[stage],
                                         kb * KBLOCK_ELEMS, static_cast<int32_t>(x_row), p.hint_x);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer: all slabs back to back =====================
            ptx::mbar_wait(qfull_bar, 0);
            ptx::tc_fence_after();
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int s = 0; s < sp.n_slabs; s++) {
                const int tiles = static_cast<int>((sp.slab_row[s + 1] - sp.slab_row[s] + STREAM_TILE_X - 1) / STREAM_TILE_X);
                for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                    ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * NQ);
                    for (int kb = 0; kb < p.num_k_blocks; kb++) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t a_addr = ptx::smem_u32(smem_x + stage * STREAM_X_STAGE_BYTES);
                        const uint32_t b_addr = ptx::smem_u32(smem_q + kb * Q_KB_BYTES);
#pragma unroll
                        for (int k = 0; k < KBLOCK_BYTES / UMMA_K_BYTES; k++) {
                            const uint64_t adesc = ptx::umma_desc_k_sw128(a_addr + k * UMMA_K_BYTES);
                            const uint64_t bdesc = ptx::umma_desc_k_sw128(b_addr + k * UMMA_K_BYTES);
                            ptx::mma_f16_ss(d_tmem, adesc, bdesc, p.idesc, (kb | k) != 0);
                        }
                        ptx::tc_commit(&empty_bar[stage]);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
                    ptx::tc_commit(&tfull_bar[as]);
                    if (++as == STREAM_ACC) { as = 0; aphase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ============ filter epilogue (thread = corpus row) + threshold refresh between slabs ============
        const int wq = warp & 3;
        const int etid = EpilogueGroup::tid();
        unsigned int* const arrivals = sp.sync;
        unsigned int* const abort_word = sp.sync + 1;
        unsigned int* const refreshed = sp.sync + 2;
        float thr_r[NQ];
#pragma unroll
        for (int j = 0; j < NQ; j++)
            thr_r[j] = j < p.nq ? __ldcg(sp.thr + j) : __int_as_float(0x7f800000);   // +inf: padding
        int as = 0;
        uint32_t aphase = 0;
        bool aborted = false;
        for (int s = 0; s < sp.n_slabs; s++) {
            const long long r0 = sp.slab_row[s], r_end = sp.slab_row[s + 1];
            const int tiles = static_cast<int>((r_end - r0 + STREAM_TILE_X - 1) / STREAM_TILE_X);
            const bool dense = s == 0 && sp.dense_first;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const long long row = r0 + static_cast<long long>(t) * STREAM_TILE_X + wq * 32 + lane;
                const bool row_ok = row < r_end;
                ptx::mbar_wait(&tfull_bar[as], aphase);
                ptx::tc_fence_after();
                const uint32_t taddr =
                    tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(as * NQ);
#pragma unroll
                for (int c = 0; c < NQ / 32; c++) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c * 32, v);
                    ptx::tmem_ld_wait();
                    if (dense) {
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (c * 32 + j < p.nq)
                                    p.cand[static_cast<long long>(c * 32 + j) * p.cap + (row - r0)] =
                                        make_key(__uint_as_float(v[j]), static_cast<uint32_t>(row));
                        }
                    } else {
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            mask |= (__uint_as_float(v[j]) > thr_r[c * 32 + j]) ? (1u << j) : 0u;
                        if (!row_ok) mask = 0;
                        while (mask) {
                            const int j = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int q = c * 32 + j;
                            const uint32_t bits = select32(v, j);
                            const int slot = atomicAdd(p.cnt + q, 1);
                            if (slot < p.cap)
                                p.cand[static_cast<long long>(q) * p.cap + slot] =
                                    make_key(__uint_as_float(bits), static_cast<uint32_t>(row));
                        }
                    }
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tempty_bar[as]);
                if (++as == STREAM_ACC) { as = 0; aphase ^= 1; }
            }
            if (s + 1 == sp.n_slabs) break;      // the refresh after the last slab belongs to finalize

            // ---- slab s is filtered by this CTA: arrive; the CTAs that own a query refresh it ----
            __threadfence();                     // this thread's list appends, device-wide
            EpilogueGroup::sync();
            if (etid == 0) red_release_gpu_add(arrivals, 1u);
            const unsigned int ep = static_cast<unsigned int>(s + 1);
            for (int q = blockIdx.x; q < p.nq; q += gridDim.x) {
                if (etid == 0 && !aborted)
                    r_warp[7] = stream_wait(arrivals, ep * gridDim.x, abort_word, sp.timeout_ns) ? 1 : 0;
                EpilogueGroup::sync();
                if (!aborted && r_warp[7] == 0) aborted = true;
                EpilogueGroup::sync();           // r_warp is reused by refresh_list
                if (!aborted) {
                    if (!(sp.flags[q] & FLAG_OVERFLOW))
                        refresh_list<EpilogueGroup>(q, sp.k, p.cap, p.cand, p.cnt, sp.kept, sp.thr, sp.eps2,
                                                    sp.flags, sp.gstats, r_keys, r_hist, r_prefix, r_krem,
                                                    r_warp, STREAM_REFRESH_KEYS);
                    __threadfence();             // compacted list, counters, threshold
                    EpilogueGroup::sync();
                }
                if (etid == 0) red_release_gpu_add(refreshed, 1u);
            }
            // ---- thresholds of slab s + 1 ----
            if (etid == 0 && !aborted)
                r_warp[7] = stream_wait(refreshed, ep * static_cast<unsigned int>(p.nq), abort_word, sp.timeout_ns) ? 1 : 0;
            EpilogueGroup::sync();
            if (!aborted && r_warp[7] == 0) aborted = true;
#pragma unroll
            for (int j = 0; j < NQ; j++)
                if (j < p.nq) thr_r[j] = __ldcg(sp.thr + j);
            EpilogueGroup::sync();               // r_warp[7] is rewritten in the next round
        }
        if (aborted || (blockIdx.x == 0 && *reinterpret_cast<volatile unsigned int*>(abort_word) != 0u)) {
            // Some wait was cut short somewhere: lists may have been appended to while they were
            // compacted.  Everything goes to the exact path (flags are read by finalize and the host).
            for (int q = etid; q < p.nq; q += EpilogueGroup::size()) atomicOr(sp.flags + q, FLAG_OVERFLOW);
            if (etid == 0)
                atomicAdd(reinterpret_cast<unsigned long long*>(sp.gstats + GS_OVERFLOW),
                          static_cast<unsigned long long>(p.nq));
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kTmemCols>(tmem_base);
    }
}

}  // namespace b2ip
