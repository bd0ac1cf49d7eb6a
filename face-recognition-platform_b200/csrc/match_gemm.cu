// K4 — tcgen05/TMEM 1:N matcher for query batches B > 4.
//
// Replaces the per-query Python loop of batch_compare_faces
// (backend/app/services/face_service.py:448-481: B x face_distance + `<=`) and the
// cosine 1:N of insightface compute_sim by ONE dense contraction S = Q . G^T on the
// 5th-gen tensor cores with the distance/threshold/top-k fused into the epilogue:
// scores never leave the SM.
//
// Operand roles: A = 128 queries (M, resident in shared memory for the whole kernel),
// B = a 256-row gallery tile (N), K = dimp in blocks of 64 halfs (one 128-byte
// swizzle span).  D[128 x 256] fp32 lives in TMEM: lane = query, column = gallery row,
// so each epilogue thread owns ONE query and scans its 256 scores per tile against a
// running cut, keeping a sorted top-KP list in registers (no cross-thread traffic).
//
// Warp roles (192 or 320 threads, one CTA per SM, persistent over its gallery tiles):
//   warp 0   TMA producer: queries once, then the gallery ring (cp.async.bulk.tensor
//            2-D, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (cta_group::1,
//            kind::f16, M=128 N=256 K=16), tcgen05.commit -> smem-slot / TMEM-stage
//            mbarriers
//   warps 2.. epilogue: tcgen05.ld 32x32b.x32 of their TMEM lane quadrant, cut test,
//            register top-KP insert; double-buffered against the MMA through two
//            256-column accumulator stages.  Four warps (one per quadrant) for long rows;
//            eight (two per quadrant, 128 columns each, two candidate lists per query
//            per CTA) for dimp <= 128, where the tensor pipe would otherwise wait for
//            the scan.
// Batches of 129..256 (any even number of 128-query groups) run as clusters of two CTAs
// that stream the SAME gallery tiles: each CTA fetches half of every tile and
// TMA-multicasts it into both CTAs' shared memory (UTMALDG.2D.MULTICAST), the smem slot
// is released in both by a multicast tcgen05.commit — half the L2 -> SM traffic.
// L2 metric: |g|^2 of each tile reaches the epilogue through a 4-slot bulk-copy ring.
// Output contract is the scan kernel's: cand[B][nlists][KP] u64 keys.
//
// Rows >= 256-d with an even number of groups use match_gemm_pair_kernel instead: the same
// pipeline with cta_group::2 (M=256) MMAs issued by the leader CTA of the pair (below).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "frp_device.cuh"
#include "frp_internal.h"

namespace frp {

namespace {

constexpr int kTileM = 128;        // queries per CTA
constexpr int kTileN = 256;        // gallery rows per MMA tile
constexpr int kBlockK = 64;        // halfs per K block (128-byte swizzle span)
constexpr int kUmmaK = 16;
// LPC ("lists per CTA", template parameter) = epilogue warps per TMEM lane quadrant: 1 -> 4
// epilogue warps that scan whole tiles, 2 -> 8 warps that scan half a tile each.
constexpr int threads_for(int lpc) { return 64 + 128 * lpc; }
constexpr int kABlockBytes = kTileM * kBlockK * 2;   // 16 KB
constexpr int kBStageBytes = kTileN * kBlockK * 2;   // 32 KB
constexpr int kMaxStages = 6;
constexpr int kMaxPairStages = 12;                  // CTA-pair kernel: half-tile (16 KB) slots
constexpr int kAccCols = kTileN;                     // fp32 accumulator columns per stage
constexpr uint32_t kTmemCols = 512;
constexpr int kNormSlots = 4;                        // L2 metric: |g|^2 tiles in flight
constexpr int kNormRingBytes = kNormSlots * kTileN * 4;

struct GemmArgs {
    const uint8_t* alive;
    const float* norm2;
    const QParams* qp;
    uint64_t* cand;
    uint32_t rows;
    uint32_t ntiles;
    int B;
    int nkb;       // K blocks (dimp / 64)
    int stages;    // gallery ring depth
    int metric;
};

// ---- PTX wrappers --------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}


// The MMA issuer runs warp-convergent: every lane executes the statement and ONE lane, chosen by
// elect.sync inside it, issues the instruction.  Under a divergent `if (lane == 0)` ptxas wraps
// each tcgen05 instruction (uniform-datapath operands) in an ELECT / PLOP3 / BRA.U.ANY loop over
// the active lanes; with 128-d rows (two K blocks per tile) that made the issuer warp, not the
// tensor pipe, the critical path.  FRP_MMA_ELECT=0 rebuilds the old form (measurement aid).
#ifndef FRP_MMA_ELECT
#define FRP_MMA_ELECT 1
#endif
#if FRP_MMA_ELECT
#define FRP_ELECT_DECL ".reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
#define FRP_ELECT_PRED "@e "
#else
#define FRP_ELECT_DECL ""
#define FRP_ELECT_PRED ""
#endif
template <int CTA_GROUP>
__device__ __forceinline__ void umma_f16_issue(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    if (CTA_GROUP == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t" FRP_ELECT_DECL
            "setp.ne.b32 p, %4, 0;\n\t"
            FRP_ELECT_PRED "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t" FRP_ELECT_DECL
            "setp.ne.b32 p, %4, 0;\n\t"
            FRP_ELECT_PRED "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// commit to this CTA's barrier (cta_group::1, no multicast)
__device__ __forceinline__ void umma_commit_issue(uint64_t* bar) {
    asm volatile(
        "{\n\t" FRP_ELECT_DECL FRP_ELECT_PRED
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
// commit multicast to the barrier at the same offset in every CTA of `mask`
template <int CTA_GROUP>
__device__ __forceinline__ void umma_commit_mc_issue(uint64_t* bar, uint16_t mask) {
    if (CTA_GROUP == 1) {
        asm volatile(
            "{\n\t" FRP_ELECT_DECL FRP_ELECT_PRED
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
            ::"r"(smem_u32(bar)), "h"(mask)
            : "memory");
    } else {
        asm volatile(
            "{\n\t" FRP_ELECT_DECL FRP_ELECT_PRED
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
            ::"r"(smem_u32(bar)), "h"(mask)
            : "memory");
    }
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
// start address >> 4 | LBO = 1 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2,
// built in two halves: the low word carries the start address (>> 4) and LBO, so stepping
// to the next K slice or ring slot is one integer add (shared memory ends below 2^18, the
// 14-bit address field cannot carry); the high word is constant.
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t lo) {
    constexpr uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

// kind::f16 instruction descriptor: D fp32, A/B fp16, both K-major, M=128, N=256
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

// ---- thread-block cluster helpers (CS = 2: the two query groups that stream the same
// gallery tiles form a cluster; each CTA fetches half of every tile from L2 and
// multicasts it into both CTAs' shared memory, halving the L2 -> SM traffic) ----------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
// Arrive on a barrier of another CTA of the cluster.  Default (cta-scope release) semantics as
// in CUTLASS' ClusterBarrier::arrive: what the arrival orders here is TMEM traffic, which the
// tcgen05 fences around it cover; a cluster-scope release costs a full memory barrier
// (SASS ERRBAR, 30 % of the epilogue's stall samples when it was tried).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}

// ---- epilogue warps (shared by the one-CTA and the CTA-pair kernels) ------------------
// list / nlist: this CTA's tile stream (tiles list, list + nlist, ...) and candidate-list slot.
template <int KP, bool IS_L2, int LPC, bool PAIR>
__device__ __forceinline__ void epilogue_warps(const GemmArgs& a, const int warp, const int lane, const int qgroup,
                                               const uint32_t list, const uint32_t nlist, const uint32_t my_tiles,
                                               const uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                               uint64_t* nfull_bar, uint64_t* nempty_bar, float* norm_ring,
                                               const uint32_t tempty_leader) {
    // ================= epilogue: one query per thread, half a tile per warp =================
    // Eight warps: warp w may only touch TMEM lane quadrant w % 4, so two warps share a
    // quadrant (= the same 32 queries) and split the tile's 256 columns.  Two warps per
    // scheduler hide each other's tcgen05.ld / shared-memory / FMNMX latencies; each
    // keeps its own top-KP list, so a CTA emits two lists per query.
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    constexpr int kEpiCols = kTileN / LPC;          // tile columns scanned by one epilogue warp
    const int half = (warp - 2) >> 2;               // which column range of every tile
    const int q = qgroup * kTileM + quad * 32 + lane;
    const bool valid_q = q < a.B;
    const float ninf = __int_as_float(0xff800000);
    griddep_wait();        // the cut comes from prep_queries_kernel (or raise_cut_kernel)
    float cut0 = valid_q ? a.qp[q].cut0 : __int_as_float(0x7f800000);
    float cut = cut0;
    float ls[KP];
    uint32_t lr[KP];
#pragma unroll
    for (int p = 0; p < KP; ++p) { ls[p] = ninf; lr[p] = 0xffffffffu; }
    const uint32_t norm_u32 = smem_u32(norm_ring);

    for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t as = i & 1;
        const uint32_t row0 = (list + i * nlist) * kTileN + half * kEpiCols;
        const uint32_t ns = i % kNormSlots;
        const uint32_t nsm = norm_u32 + (ns * kTileN + half * kEpiCols) * 4;
        if (IS_L2) mbar_wait(&nfull_bar[ns], (i / kNormSlots) & 1);
        mbar_wait(&tfull_bar[as], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kAccCols + half * kEpiCols;
#pragma unroll 1
        for (int c = 0; c < kEpiCols; c += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + c, v);
            tmem_ld_wait();
            float sv[32];
            if (IS_L2) {
                // s = 2 g.q - |g|^2 ; every lane reads the same shared-memory words (broadcast)
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 n4 = lds_f4(nsm + (c + j) * 4);
                    sv[j] = fmaf(2.f, __uint_as_float(v[j]), -n4.x);
                    sv[j + 1] = fmaf(2.f, __uint_as_float(v[j + 1]), -n4.y);
                    sv[j + 2] = fmaf(2.f, __uint_as_float(v[j + 2]), -n4.z);
                    sv[j + 3] = fmaf(2.f, __uint_as_float(v[j + 3]), -n4.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) sv[j] = __uint_as_float(v[j]);
            }
            // one branch per 32 scores: a per-element branch costs a convergence
            // barrier (BSSY/BSYNC) each and made the epilogue the bottleneck.  Four
            // independent max chains keep the dependent-issue depth at 8 + 2.
            float m0 = sv[0], m1 = sv[1], m2 = sv[2], m3 = sv[3];
#pragma unroll
            for (int j = 4; j < 32; j += 4) {
                m0 = fmaxf(m0, sv[j]);
                m1 = fmaxf(m1, sv[j + 1]);
                m2 = fmaxf(m2, sv[j + 2]);
                m3 = fmaxf(m3, sv[j + 3]);
            }
            const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            if (mx >= cut) {
                // Rare path, taken by the whole warp when any lane has a candidate: test the four
                // max chains first so that only the 8 scores of a chain that can hold a candidate
                // are walked (a gallery with real matches takes this path for ~10 % of the
                // 32-score groups; walking all 32 scores each time cost 40 % of the kernel).
                // Insertion order is chain-major, not column-major: ties at the tail of a full
                // list may therefore drop the lower row, which the certificate (strict
                // s_kth > L + eps) turns into a fallback, never into a wrong answer.
                const float mch[4] = {m0, m1, m2, m3};
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    if (mch[ch] >= cut) {
#pragma unroll
                        for (int j = ch; j < 32; j += 4) {
                            const float s = sv[j];
                            const uint32_t row = row0 + c + j;
                            if (s >= cut && row < a.rows) {
                                if (a.alive[row]) {
                                    float cs = s;
                                    uint32_t cr = row;
#pragma unroll
                                    for (int p = 0; p < KP; ++p) {
                                        const bool gt = cs > ls[p];
                                        const float ts = gt ? ls[p] : cs;
                                        const uint32_t tr = gt ? lr[p] : cr;
                                        ls[p] = gt ? cs : ls[p];
                                        lr[p] = gt ? cr : lr[p];
                                        cs = ts;
                                        cr = tr;
                                    }
                                    cut = fmaxf(cut0, ls[KP - 1]);
                                }
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            // PAIR: the MMA issuer lives in the leader CTA and waits for both CTAs' scans
            if (PAIR) mbar_arrive_cluster(tempty_leader + as * 8);
            else mbar_arrive(&tempty_bar[as]);
            if (IS_L2) mbar_arrive(&nempty_bar[ns]);
        }
    }
    if (valid_q) {
        uint64_t* out = a.cand + ((size_t)q * (nlist * LPC) + list * LPC + half) * KP;
#pragma unroll
        for (int p = 0; p < KP; ++p) out[p] = (lr[p] != 0xffffffffu) ? make_key(ls[p], lr[p]) : 0ull;
    }
}

template <int KP, bool IS_L2, int CS, int LPC>
__global__ void __launch_bounds__(threads_for(LPC), 1)
match_gemm_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                  const __grid_constant__ CUtensorMap tmap_gh, const GemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B operands
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)a.nkb * kABlockBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)a.stages * kBStageBytes);
    uint64_t* full_bar = bars;                       // [kMaxStages]
    uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
    uint64_t* tfull_bar = bars + 2 * kMaxStages;     // [2]
    uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;  // [2]
    uint64_t* afull_bar = bars + 2 * kMaxStages + 4;   // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);
    // L2 metric: |g|^2 of the tile's 256 rows travels through its own 4-deep ring (1 KB bulk
    // copies) so the epilogue reads it from shared memory (broadcast) instead of 256 global
    // loads per tile per warp
    uint64_t* nfull_bar = bars + 2 * kMaxStages + 6;    // [kNormSlots]
    uint64_t* nempty_bar = nfull_bar + kNormSlots;      // [kNormSlots]
    float* norm_ring = reinterpret_cast<float*>(bars + 32);   // [kNormSlots][kTileN], 256 B after the barriers

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qgroup = blockIdx.y;
    const uint32_t my_tiles = (a.ntiles > blockIdx.x) ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_g);
        prefetch_tmap(&tmap_gh);
        // a smem slot is free when the MMAs of EVERY CTA of the cluster have retired (the
        // peer multicasts into this CTA's slot too)
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CS); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4 * LPC); }
        mbar_init(afull_bar, 1);
        for (int s = 0; s < kNormSlots; ++s) { mbar_init(&nfull_bar[s], 1); mbar_init(&nempty_bar[s], 4 * LPC); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();      // peer barriers are initialised before any multicast lands
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t crank = (CS > 1) ? cluster_ctarank() : 0u;
    griddep_launch();      // the re-rank launch is staged behind this kernel (it waits for its end)

    if (warp == 0) {
        // ================= TMA producer =================
        // The gallery ring is filled first: this kernel may have started while prep_queries_kernel
        // is still running (programmatic dependent launch), and the gallery does not depend on it.
        // The queries are fetched after the first `stages` K blocks, behind griddep_wait().
        auto load_queries = [&]() {
            if (lane == 0) {
                griddep_wait();
                mbar_expect_tx(afull_bar, (uint32_t)a.nkb * kABlockBytes);
                for (int kb = 0; kb < a.nkb; ++kb)
                    tma_load_2d(&tmap_q, afull_bar, smem_a + (size_t)kb * kABlockBytes, kb * kBlockK, qgroup * kTileM);
            }
            __syncwarp();
        };
        bool q_pending = true;
        int issued = 0;
        int s = 0;
        uint32_t ph = 0;
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const int row0 = (int)((blockIdx.x + i * gridDim.x) * kTileN);
            if (IS_L2) {
                const uint32_t ns = i % kNormSlots;
                mbar_wait(&nempty_bar[ns], ((i / kNormSlots) & 1) ^ 1);     // epilogue released the slot
                if (lane == 0) {
                    mbar_expect_tx(&nfull_bar[ns], kTileN * 4);
                    // norm2 is allocated with a tile of slack, so a full tile may always be read
                    bulk_g2s(norm_ring + ns * kTileN, a.norm2 + row0, kTileN * 4, &nfull_bar[ns]);
                }
                __syncwarp();
            }
            for (int kb = 0; kb < a.nkb; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(&full_bar[s], kBStageBytes);
                    if (CS > 1) {
                        // this CTA's half of the tile (128 rows), delivered to both CTAs
                        tma_load_2d_mc(&tmap_gh, &full_bar[s],
                                       smem_b + (size_t)s * kBStageBytes + (size_t)crank * (kBStageBytes / 2),
                                       kb * kBlockK, row0 + (int)crank * (kTileN / 2), (uint16_t)0x3);
                    } else {
                        tma_load_2d(&tmap_g, &full_bar[s], smem_b + (size_t)s * kBStageBytes, kb * kBlockK, row0);
                    }
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1; }
                if (q_pending && ++issued == a.stages) { load_queries(); q_pending = false; }
            }
        }
        if (q_pending) load_queries();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        mbar_wait(afull_bar, 0);
        tc_fence_after();
        int s = 0;
        uint32_t ph = 0;
        const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem_a)), b_lo0 = smem_desc_lo(smem_u32(smem_b));
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const uint32_t as = i & 1;
            mbar_wait(&tempty_bar[as], ((i >> 1) & 1) ^ 1);     // epilogue drained this TMEM stage
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * kAccCols;
            for (int kb = 0; kb < a.nkb; ++kb) {
                mbar_wait(&full_bar[s], ph);                    // TMA landed this gallery K block
                tc_fence_after();
#if !FRP_MMA_ELECT
                if (lane == 0)
#endif
                {
                    const uint32_t a_lo = a_lo0 + (uint32_t)kb * (kABlockBytes >> 4);
                    const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBStageBytes >> 4);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_f16_issue<1>(tmem_d, smem_desc(a_lo + k * (kUmmaK * 2 >> 4)),
                                          smem_desc(b_lo + k * (kUmmaK * 2 >> 4)), kIdesc, (kb | k) != 0 ? 1u : 0u);
                    // frees the smem slot (in every CTA of the cluster) when the MMAs retire
                    if (CS > 1) umma_commit_mc_issue<1>(&empty_bar[s], (uint16_t)0x3);
                    else umma_commit_issue(&empty_bar[s]);
                    if (kb == a.nkb - 1) umma_commit_issue(&tfull_bar[as]);   // accumulator ready
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        epilogue_warps<KP, IS_L2, LPC, false>(a, warp, lane, qgroup, blockIdx.x, gridDim.x, my_tiles, tmem_base,
                                              tfull_bar, tempty_bar, nfull_bar, nempty_bar, norm_ring, 0u);
    }

    tc_fence_before();
    __syncthreads();
    // no CTA of the cluster may exit while its peer can still multicast data or barrier
    // arrivals into its shared memory
    if (CS > 1) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- CTA-pair variant (cta_group::2) ------------------------------------------------
// The two 128-query groups of a 256-query batch run as one CTA pair issuing M=256 MMAs:
// each CTA keeps its own 128 queries (A) and stages only ITS half of every gallery tile
// (128 of the 256 rows, B) — the tensor cores of both SMs read both halves, so the
// shared-memory traffic per MMA halves and no multicast copy is needed.  Plumbing
// (as CUTLASS' 2-SM kernels do it): both CTAs' TMA loads complete on the LEADER's full
// barrier (cp.async.bulk.tensor ... cta_group::2), only the leader's warp 1 issues
// tcgen05.mma.cta_group::2, its commits are multicast to both CTAs' barriers, and the
// peer's epilogue warps release the accumulator stage on the leader's barrier.
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// dst in this CTA's shared memory, completion bytes on the barrier at cluster address bar_caddr
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t bar_caddr, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_caddr), "r"(c0), "r"(c1)
        : "memory");
}

constexpr int kBHalfBytes = kBStageBytes / 2;      // 128 gallery rows x 64 halfs
// kind::f16 instruction descriptor for the pair: M = 256 (128 per CTA), N = 256
constexpr uint32_t kIdescPair = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)((2 * kTileM) >> 4) << 24);

template <int KP, bool IS_L2, int LPC>
__global__ void __launch_bounds__(threads_for(LPC), 1)
match_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_gh,
                       const GemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)a.nkb * kABlockBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)a.stages * kBHalfBytes);
    uint64_t* full_bar = bars;                       // [kMaxPairStages]   (used in the leader)
    uint64_t* empty_bar = bars + kMaxPairStages;     // [kMaxPairStages]
    uint64_t* tfull_bar = bars + 2 * kMaxPairStages;       // [2]
    uint64_t* tempty_bar = bars + 2 * kMaxPairStages + 2;  // [2]          (used in the leader)
    uint64_t* afull_bar = bars + 2 * kMaxPairStages + 4;   // [1]          (used in the leader)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxPairStages + 5);
    uint64_t* nfull_bar = bars + 2 * kMaxPairStages + 6;   // [kNormSlots]
    uint64_t* nempty_bar = nfull_bar + kNormSlots;         // [kNormSlots]
    float* norm_ring = reinterpret_cast<float*>(bars + 2 * kMaxPairStages + 6 + 2 * kNormSlots + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();        // 0 = leader
    const int qgroup = blockIdx.y * 2 + (int)crank;
    const uint32_t list = blockIdx.x >> 1, nlist = gridDim.x >> 1;
    const uint32_t my_tiles = (a.ntiles > list) ? (a.ntiles - list + nlist - 1) / nlist : 0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_gh);
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * 4 * LPC); }
        mbar_init(afull_bar, 1);
        for (int s = 0; s < kNormSlots; ++s) { mbar_init(&nfull_bar[s], 1); mbar_init(&nempty_bar[s], 4 * LPC); }
        mbar_fence_init();
    }
    __syncthreads();
    cluster_sync_all();                 // both CTAs' barriers exist before anything is signalled
    if (warp == 1) tmem_alloc2(tmem_ptr, kTmemCols);     // the same warp in both CTAs
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_launch();      // the re-rank launch is staged behind this kernel (it waits for its end)
    // cluster addresses of the leader's barriers
    const uint32_t full_leader = mapa_u32(smem_u32(full_bar), 0);
    const uint32_t afull_leader = mapa_u32(smem_u32(afull_bar), 0);
    const uint32_t tempty_leader = mapa_u32(smem_u32(tempty_bar), 0);

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        // gallery ring first, queries behind griddep_wait() (see match_gemm_kernel)
        auto load_queries = [&]() {
            if (lane == 0) {
                griddep_wait();
                if (crank == 0) mbar_expect_tx(afull_bar, 2u * (uint32_t)a.nkb * kABlockBytes);
                for (int kb = 0; kb < a.nkb; ++kb)
                    tma_load_2d_2sm(&tmap_q, afull_leader, smem_a + (size_t)kb * kABlockBytes, kb * kBlockK,
                                    qgroup * kTileM);
            }
            __syncwarp();
        };
        bool q_pending = true;
        int issued = 0;
        int s = 0;
        uint32_t ph = 0;
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const int row0 = (int)((list + i * nlist) * kTileN);
            if (IS_L2) {
                const uint32_t ns = i % kNormSlots;
                mbar_wait(&nempty_bar[ns], ((i / kNormSlots) & 1) ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(&nfull_bar[ns], kTileN * 4);
                    bulk_g2s(norm_ring + ns * kTileN, a.norm2 + row0, kTileN * 4, &nfull_bar[ns]);
                }
                __syncwarp();
            }
            for (int kb = 0; kb < a.nkb; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);      // the pair's MMAs on this slot have retired
                if (lane == 0) {
                    if (crank == 0) mbar_expect_tx(&full_bar[s], kBStageBytes);     // both halves
                    tma_load_2d_2sm(&tmap_gh, full_leader + (uint32_t)s * 8, smem_b + (size_t)s * kBHalfBytes,
                                    kb * kBlockK, row0 + (int)crank * (kTileN / 2));
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1; }
                if (q_pending && ++issued == a.stages) { load_queries(); q_pending = false; }
            }
        }
        if (q_pending) load_queries();
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (crank == 0) {
            mbar_wait(afull_bar, 0);
            tc_fence_after();
            int s = 0;
            uint32_t ph = 0;
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem_a)), b_lo0 = smem_desc_lo(smem_u32(smem_b));
            for (uint32_t i = 0; i < my_tiles; ++i) {
                const uint32_t as = i & 1;
                mbar_wait(&tempty_bar[as], ((i >> 1) & 1) ^ 1);     // both CTAs drained this TMEM stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * kAccCols;
                for (int kb = 0; kb < a.nkb; ++kb) {
                    mbar_wait(&full_bar[s], ph);                    // both halves of the K block landed
                    tc_fence_after();
#if !FRP_MMA_ELECT
                    if (lane == 0)
#endif
                    {
                        const uint32_t a_lo = a_lo0 + (uint32_t)kb * (kABlockBytes >> 4);
                        const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBHalfBytes >> 4);
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k)
                            umma_f16_issue<2>(tmem_d, smem_desc(a_lo + k * (kUmmaK * 2 >> 4)),
                                              smem_desc(b_lo + k * (kUmmaK * 2 >> 4)), kIdescPair,
                                              (kb | k) != 0 ? 1u : 0u);
                        umma_commit_mc_issue<2>(&empty_bar[s], (uint16_t)0x3);              // slot free in both CTAs
                        if (kb == a.nkb - 1) umma_commit_mc_issue<2>(&tfull_bar[as], (uint16_t)0x3);   // accumulators ready
                    }
                    __syncwarp();
                    if (++s == a.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        epilogue_warps<KP, IS_L2, LPC, true>(a, warp, lane, qgroup, list, nlist, my_tiles, tmem_base, tfull_bar,
                                             tempty_bar, nfull_bar, nempty_bar, norm_ring, tempty_leader);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // the peer may still signal this CTA's barriers / read its shared memory
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, kTmemCols);
    }
}

// ---- host side -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
    });
    return fn;
}

// 2-D fp16 row-major [rows][dimp] tensor, box = [box_rows][64], 128-byte swizzle
int encode_map(CUtensorMap* map, const void* base, uint64_t rows, int dimp, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return FRP_E_CUDA; }
    cuuint64_t gdim[2] = {(cuuint64_t)dimp, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dimp * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return FRP_E_CUDA; }
    return FRP_OK;
}

int pick_stages(int nkb, bool l2) {
    // barriers + alignment slack (+ the |g|^2 ring of the L2 epilogue)
    const int budget = 227 * 1024 - 2048 - nkb * kABlockBytes - (l2 ? kNormRingBytes : 0);
    int st = budget / kBStageBytes;
    if (st > kMaxStages) st = kMaxStages;
    return st;
}

}  // namespace

void* tensormap_encode_fn() { return reinterpret_cast<void*>(get_encode_fn()); }

// SMs the prefilter leaves free (frp_set_reserved_sms): its CTAs take a whole SM's shared memory,
// so a collective that should overlap it (the all-gather + merge of the previous step on another
// stream) has nowhere to run unless a couple of SMs stay out of the grid
std::atomic<int> g_reserved_sms{0};

}  // namespace frp

extern "C" int frp_set_reserved_sms(int n) {
    if (n < 0 || n > 64) { frp::set_error("frp_set_reserved_sms: %d out of range [0, 64]", n); return FRP_E_INVALID; }
    frp::g_reserved_sms.store(n);
    return FRP_OK;
}

namespace frp {

// FRP_GEMM_NO_CLUSTER=1 disables the 2-CTA multicast variant (A/B measurements)
static const bool g_gemm_no_cluster = []() {
    const char* e = getenv("FRP_GEMM_NO_CLUSTER");
    return e && e[0] == '1';
}();

// The cta_group::2 CTA-pair kernel serves even group counts of rows >= 256-d (4 M rows, batch
// 256: 512-d 0.822 vs 0.871 ms, 256-d 0.417 vs 0.423 ms); 128-d rows are paced by the
// epilogue and run the same or slightly faster on the multicast kernel (0.229 vs 0.232 ms).  FRP_GEMM_PAIR=0|1 forces it
// off / on for every dimension (A/B measurements).
static const int g_gemm_pair = []() {
    const char* e = getenv("FRP_GEMM_PAIR");
    return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1;
}();
static bool use_pair_kernel(const frp_gallery* g) { return g_gemm_pair < 0 ? g->dimp >= 256 : g_gemm_pair == 1; }

bool gemm_supported(const frp_gallery* g) {
    return (g->dimp % kBlockK) == 0 && g->dimp <= 512 && g->rows > 0 && g->rows < (1ull << 31) &&
           g->tmap_valid;
}

static int gemm_groups(int B) { return (B + kTileM - 1) / kTileM; }

// Short rows (dimp <= 128) leave only ~1 000 tensor-pipe cycles per tile, less than one warp
// per scheduler needs for the 256-column scan: use two epilogue warps per quadrant there.
// Long rows hide the scan behind the MMAs already; the extra warps only cost power.
// FRP_GEMM_LPC=1|2 overrides (A/B measurements).
static int gemm_lpc(const frp_gallery* g) {
    static const int forced = []() {
        const char* e = getenv("FRP_GEMM_LPC");
        return (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
    }();
    if (forced) return forced;
    return g->dimp <= 128 ? 2 : 1;
}

int gemm_num_lists(const frp_gallery* g, int B) {
    const int groups = gemm_groups(B);
    int sms = g->sm_count - g_reserved_sms.load(std::memory_order_relaxed);
    if (sms < groups) sms = groups;
    int per = sms / groups;
    if (per < 1) per = 1;
    const uint64_t ntiles = (g->rows + kTileN - 1) / kTileN;
    if ((uint64_t)per > ntiles) per = (int)ntiles;
    return per * gemm_lpc(g);
}

// Descriptors of the fp16 scan copy over its whole CAPACITY: they only change when the buffers
// are reallocated (unique lock), so a match never encodes a gallery map.  Rows in
// [rows, capacity) are zero (alloc_buffers clears the copy) and masked by `row < rows`.
void gemm_refresh_maps(frp_gallery* g) {
    g->tmap_valid = false;
    if ((g->dimp % kBlockK) != 0 || g->dimp > 512 || !g->g16 || !get_encode_fn()) return;
    CUtensorMap tg, tgh;
    if (encode_map(&tg, g->g16, g->capacity, g->dimp, kTileN)) return;
    if (encode_map(&tgh, g->g16, g->capacity, g->dimp, kTileN / 2)) return;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    memcpy(g->tmap_g, &tg, sizeof(tg));
    memcpy(g->tmap_gh, &tgh, sizeof(tgh));
    g->tmap_valid = true;
}

int encode_query_map(void* tmap_out, const __half* q16, int groups, int dimp) {
    return encode_map(reinterpret_cast<CUtensorMap*>(tmap_out), q16, (uint64_t)groups * kTileM, dimp, kTileM);
}

int launch_match_gemm(frp_gallery* g, const void* tmap_q, const QParams* qp, int B, int KP, uint64_t* cand,
                      int nlists, cudaStream_t st, uint64_t rows_limit) {
    // rows_limit > 0: scan only the first rows_limit rows (the sampling pass of pure top-k)
    const uint64_t rows = (rows_limit && rows_limit < g->rows) ? rows_limit : g->rows;
    const int groups = gemm_groups(B);
    const int nkb = g->dimp / kBlockK;
    const bool l2 = (g->metric == FRP_METRIC_L2);
    const int stages = pick_stages(nkb, l2);
    if (stages < 2) { set_error("match_gemm: not enough shared memory for dimp=%d", g->dimp); return FRP_E_UNSUPPORTED; }
    if (!g->tmap_valid) { set_error("match_gemm: gallery tensor maps unavailable"); return FRP_E_UNSUPPORTED; }
    const CUtensorMap& tq = *reinterpret_cast<const CUtensorMap*>(tmap_q);
    const CUtensorMap& tg = *reinterpret_cast<const CUtensorMap*>(g->tmap_g);
    const CUtensorMap& tgh = *reinterpret_cast<const CUtensorMap*>(g->tmap_gh);
    int rc = FRP_OK;
    // an even number of 128-query groups runs as clusters of two CTAs sharing every gallery tile
    const int cs = (groups % 2 == 0 && !g_gemm_no_cluster) ? 2 : 1;
    GemmArgs a;
    a.alive = g->alive;
    a.norm2 = g->norm2;
    a.qp = qp;
    a.cand = cand;
    a.rows = (uint32_t)rows;
    a.ntiles = (uint32_t)((rows + kTileN - 1) / kTileN);
    a.B = B;
    a.nkb = nkb;
    a.stages = stages;
    a.metric = g->metric;
    const size_t smem = 1024 + (size_t)nkb * kABlockBytes + (size_t)stages * kBStageBytes + 256 +
                        (l2 ? kNormRingBytes : 0);
    const int lpc = gemm_lpc(g);
    dim3 grid(nlists / lpc, groups);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads_for(lpc));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (unsigned)cs;
    attr[0].val.clusterDim.z = 1;
    // programmatic dependent of the kernel before it in the stream (prep_queries / raise_cut)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t le = cudaSuccess;
    if (cs == 2 && use_pair_kernel(g)) {
        // CTA-pair kernel: cluster (2,1,1) = the two query groups of one 256-query slab
        int pst = (227 * 1024 - 2048 - nkb * kABlockBytes - 512 - (l2 ? kNormRingBytes : 0)) / kBHalfBytes;
        if (pst > kMaxPairStages) pst = kMaxPairStages;
        if (pst < 2) { set_error("match_gemm: not enough shared memory for dimp=%d", g->dimp); return FRP_E_UNSUPPORTED; }
        a.stages = pst;
        const size_t psmem = 1024 + (size_t)nkb * kABlockBytes + (size_t)pst * kBHalfBytes + 512 +
                             (l2 ? kNormRingBytes : 0);
        cfg.gridDim = dim3(2 * (nlists / lpc), groups / 2);
        cfg.dynamicSmemBytes = psmem;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
#define FRP_PAIR_LAUNCH(KP_, L2_, LPC_)                                                                  \
    {                                                                                                    \
        configure_func_smem((const void*)match_gemm_pair_kernel<KP_, L2_, LPC_>, g->device, psmem, &rc); \
        if (rc) return rc;                                                                               \
        le = cudaLaunchKernelEx(&cfg, match_gemm_pair_kernel<KP_, L2_, LPC_>, tq, tgh, a);               \
    }
        if (KP != 16 && KP != 32) { set_error("match_gemm: unsupported KP=%d", KP); return FRP_E_UNSUPPORTED; }
        const int pkey = (KP == 32 ? 4 : 0) | (l2 ? 2 : 0) | (lpc == 2 ? 1 : 0);
        switch (pkey) {
            case 0: FRP_PAIR_LAUNCH(16, false, 1) break;
            case 1: FRP_PAIR_LAUNCH(16, false, 2) break;
            case 2: FRP_PAIR_LAUNCH(16, true, 1) break;
            case 3: FRP_PAIR_LAUNCH(16, true, 2) break;
            case 4: FRP_PAIR_LAUNCH(32, false, 1) break;
            case 5: FRP_PAIR_LAUNCH(32, false, 2) break;
            case 6: FRP_PAIR_LAUNCH(32, true, 1) break;
            default: FRP_PAIR_LAUNCH(32, true, 2) break;
        }
#undef FRP_PAIR_LAUNCH
        if (le != cudaSuccess) {
            set_error("match_gemm (pair) launch failed: %s", cudaGetErrorString(le));
            cudaGetLastError();
            return FRP_E_CUDA;
        }
        FRP_CHECK_LAUNCH();
        return FRP_OK;
    }
#define FRP_GEMM_LAUNCH(KP_, L2_, CS_, LPC_)                                                             \
    {                                                                                                    \
        configure_func_smem((const void*)match_gemm_kernel<KP_, L2_, CS_, LPC_>, g->device, smem, &rc);  \
        if (rc) return rc;                                                                               \
        le = cudaLaunchKernelEx(&cfg, match_gemm_kernel<KP_, L2_, CS_, LPC_>, tq, tg, tgh, a);           \
    }
#define FRP_GEMM_CASES(LPC_)                                 \
    switch (key) {                                           \
        case 0: FRP_GEMM_LAUNCH(16, false, 1, LPC_) break;   \
        case 1: FRP_GEMM_LAUNCH(16, false, 2, LPC_) break;   \
        case 2: FRP_GEMM_LAUNCH(16, true, 1, LPC_) break;    \
        case 3: FRP_GEMM_LAUNCH(16, true, 2, LPC_) break;    \
        case 4: FRP_GEMM_LAUNCH(32, false, 1, LPC_) break;   \
        case 5: FRP_GEMM_LAUNCH(32, false, 2, LPC_) break;   \
        case 6: FRP_GEMM_LAUNCH(32, true, 1, LPC_) break;    \
        default: FRP_GEMM_LAUNCH(32, true, 2, LPC_) break;   \
    }
    const int key = (KP == 32 ? 4 : 0) | (l2 ? 2 : 0) | (cs == 2 ? 1 : 0);
    if (KP != 16 && KP != 32) { set_error("match_gemm: unsupported KP=%d", KP); return FRP_E_UNSUPPORTED; }
    if (lpc == 2) FRP_GEMM_CASES(2) else FRP_GEMM_CASES(1)
#undef FRP_GEMM_CASES
#undef FRP_GEMM_LAUNCH
    if (le != cudaSuccess) {
        set_error("match_gemm launch failed: %s", cudaGetErrorString(le));
        cudaGetLastError();
        return FRP_E_CUDA;
    }
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

}  // namespace frp
