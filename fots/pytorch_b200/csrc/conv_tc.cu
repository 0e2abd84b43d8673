// conv_tc.cu -- stride-1 convolution of channels-last bf16 activations as an implicit GEMM on the sm_100a
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA), with bias / leaky-ReLU fused in the
// epilogue.  It replaces the cuDNN calls for the dense 3x3 convolutions of the consumer and feeder networks
// (tools/models.py:334-379 forward_ocr conv5..conv10_s; :142-166 BasicBlockIn; :257-264 layer0_1), where the
// path really is a dense contraction.
//
// GEMM view:  D[pixel, cout] = sum_{r,s,cin} X[n, h+r-pad_h, w+s-pad_w, cin] * Wt[cout, r, s, cin]
//   M = N*Ho*Wo output pixels, tile = 128 pixels shaped TN x TH x TW (a box of the NHWC tensor)
//   N = Cout, tile BN in {64,128,256};   K = R*S*Cin walked as (tap, 64-channel chunk) blocks.
// No im2col buffer exists anywhere: for one k-block the A operand is the input box shifted by the tap, fetched by
// ONE 4-D TMA load whose out-of-bounds rows/columns are zero-filled by the hardware (that is the padding), landing
// in shared memory as a K-major 128x64 tile in the 128-byte swizzle the UMMA descriptor expects.
//
// CTA = 10 warps: warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, warps 2-9 = epilogue (TMEM -> registers
// -> bias/activation -> bf16 -> swizzled shared block -> TMA store).  Persistent: one CTA per SM walks the tile
// list; the accumulator is double-buffered in TMEM (2 x BN columns) and the producer runs ahead across tile
// boundaries, so the epilogue of tile i and the pipeline fill of tile i+1 hide behind the MMAs.
// Tile shapes: 128 pixels x 256 channels; two 128-pixel sub-tiles x 128 or 64 channels; and, for 256-wide tiles with
// enough work, CTA PAIRS (tcgen05 cta_group::2, clusters of two): 256 pixels x 256 channels per pair with half of the
// weight tile staged by each CTA (see the PAIR notes at the kernel).  Optional epilogue: per-image per-channel sum and
// sum of squares of the stored values (the statistics pass of the InstanceNorm that follows).
// Also used by CRNN.to_b200 (tools/models.py:853-909) with the BatchNorms folded into weights and bias.
#include "../../../include/fots_b200_pipeline.h"
#include "pdl.cuh"
#include <cuda.h>
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <mutex>

namespace {

constexpr int BM = 128;          // pixels per tile = UMMA M (accumulator row i lives in TMEM lane i)
constexpr int BK = 64;           // channels per k-block: 64 bf16 = one 128-byte swizzle row
constexpr int UK = 16;           // K of one tcgen05.mma.kind::f16
constexpr int kThreads = 320;       // producer warp, MMA warp, 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr uint32_t kABytes = BM * BK * 2;

struct ConvParams {
    int n_tiles_w, n_tiles_h, n_tiles_n;   // tile grid over (Wo, Ho, N)
    int tw, th, tn;                        // 128-pixel sub-tile box (tw*th*tn == 128)
    int cta_h, cta_n;                      // CTA tile extent in h and n: MT sub-tiles stacked along h (tn == 1) or n
    int sub_h, sub_n;                      // offset of sub-tile j from sub-tile j-1
    int cin_chunks;                        // Cin / 64
    int S;                                 // filter width (taps = R*S)
    int taps;
    int pad_h, pad_w;
    int stride;                            // 1 or 2 (both spatial dims): input pixel = stride * output pixel + tap - pad
    int cout_tiles;                        // Cout / BN
    int chunk;                             // consecutive tiles per CTA visit
    float slope;                           // leaky slope; 1 = identity
    const float* bias;                     // [Cout] or nullptr
    double* stats;                         // [N, Cout, 2] (sum, sum of squares) of the bf16 output per image, or nullptr
    int N, Ho, Wo, Cout;                   // output extent (to mask the rows of ragged tiles out of the statistics)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spin on the phase parity.  A barrier that never flips would hang the device until the driver's watchdog; trap
// after ~2 s instead so a broken pipeline shows up as a launch failure.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((spin & 0xfff) == 0xfff) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.  Advancing K by 16 bf16 inside the swizzle row is
// +32 bytes on the start address (+2 in the 16-byte units of the field).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Same, for an operand whose 8-row groups are `sbo` bytes apart and whose first row is NOT on a 1024-byte boundary of the
// swizzle pattern (single-box halo mode: the start is a whole number of 128-byte rows into the box the TMA wrote).
// MEASURED on B200 (tools/halo_box_check.py): the tensor core applies the 128-byte swizzle to the absolute shared-memory
// address bits, exactly as the TMA did when it wrote the box, so the descriptor needs NO base_offset (bits 49-51 = 0);
// setting base_offset to the start's row phase gives wrong results.
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster act as one 256-row MMA; only the even CTA issues ----
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;       // clears the CTA-rank bit of a shared::cluster address -> the even CTA of the pair
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & kPeerMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_even_cta(uint32_t bar) {      // plain arrival on the even CTA's copy of `bar`
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {          // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}

// PAIR: the kernel is launched in clusters of two CTAs that act as one 256-pixel x BN tile (tcgen05 cta_group::2):
// each CTA stages its own 128 pixels of A and HALF of the weight tile (BN/2 rows), the even CTA issues M = 256 MMAs that
// read both shared memories and write both TMEMs, and both CTAs run their own epilogue.  A third fewer operand bytes
// per MMA cycle and per stage -> six stages instead of four in the same shared memory.
// HALO (3x3, stride 1, pad 1 on large maps): a pipeline stage is not one filter tap but one COLUMN of taps (s fixed,
// r = 0..2) of a 64-channel chunk.  The CTA tile is 8 pixels wide and MT x 16 pixels tall; ONE TMA box of 8 x (MT*16 + 2)
// pixels -- the tile plus one halo row above and below, shifted by s - 1 pixels -- lands in shared memory as
// [row][8 pixels][128 B], i.e. every image row is exactly one 1024-byte swizzle atom, so the A operand of tap (r, s) for
// sub-tile j is the SAME buffer at byte offset (16 j + r) * 1024: three taps are fed from one load.  Operand bytes per
// MMA cycle drop 2.0x (64-wide) / 1.7x (128-wide cout tiles), which is what bounds these shapes: with 64 output
// channels the per-tap form needs 160 B/clk/SM from L2 against ~42 available (26 % tensor utilisation, as measured).
// WRES (with HALO, Cin == 64, Cout == BN): the whole 3 x 3 x 64 x BN weight block (72 KB at BN = 64) is loaded ONCE per
// CTA and stays in shared memory for every tile the persistent CTA walks; the ring then carries only the A halos.
template <int MT, int BN, int STAGES, bool PAIR = false, bool HALO = false, bool WRES = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_y, const ConvParams P) {
    pdl::trigger();          // programmatic dependent launch (pdl.cuh): barrier / TMEM set-up below overlaps the previous kernel's tail
    static_assert(!PAIR || MT == 1, "pair mode: one 128-pixel sub-tile per CTA");
    static_assert(!(PAIR && HALO), "halo reuse is a single-CTA mode");
    constexpr uint32_t kBBytes = (PAIR ? BN / 2 : BN) * BK * 2;
    constexpr uint32_t kAHalo = (MT * 16 + 2) * 1024;            // HALO: (MT*16 + 2) image rows of 8 pixels x 128 B
    static_assert(!WRES || HALO, "resident weights are a halo-mode option");
    constexpr uint32_t kWRes = WRES ? 9 * kBBytes : 0;           // resident weight block in front of the ring
    constexpr uint32_t kABox = (((MT * 16 + 2) * 10 * 128) + 1023) / 1024 * 1024;   // single-box halo: (MT*16 + 2) rows of 10 pixels
    constexpr uint32_t kStageBytes = WRES ? (STAGES == 2 ? kABox : kAHalo) : HALO ? kAHalo + 3 * kBBytes : MT * kABytes + kBBytes;   // MT pixel sub-tiles share one weight tile
    constexpr bool kBox = WRES && STAGES == 2;                   // the 2-stage WRES instantiation is the single-box form
    constexpr uint32_t kOutBlk = BM * 128;                       // staging block: 128 pixels x 64 channels of bf16
    constexpr uint32_t kAccCols = MT * BN;                       // one accumulator set: MT sub-tiles x BN columns
    constexpr uint32_t kTmemCols = 2 * kAccCols;                 // two sets: epilogue of tile i overlaps MMA of i+1
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns: power of two <= 512");
    // instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t wres = (raw + 1023u) & ~1023u;                 // swizzle-128B tiles need 1024-byte alignment
    const uint32_t base = wres + kWRes;                           // ring of pipeline stages
    uint8_t* const base_ptr = smem_raw + (base - raw);
    const uint32_t out_stage = base + STAGES * kStageBytes;       // 2 staging blocks for the TMA stores
    uint8_t* const out_ptr = base_ptr + STAGES * kStageBytes;
    const uint32_t bars = out_stage + 2 * kOutBlk;                // full[STAGES], empty[STAGES], tfull[2], tempty[2], tmem slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t wres_bar = bars + 8u * (2 * STAGES + 4);
    volatile uint32_t* const tmem_slot =
        reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * kStageBytes + 2 * kOutBlk + 8 * (2 * STAGES + 5));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = kBox ? 1 : HALO ? 3 * P.cin_chunks : P.taps * P.cin_chunks;      // HALO: k-block = (64-channel chunk, tap column s)
    const int num_mtiles = P.n_tiles_w * P.n_tiles_h * P.n_tiles_n;
    const int num_tiles = P.cout_tiles * (PAIR ? (num_mtiles + 1) / 2 : num_mtiles);
    // This CTA's tiles: runs of P.chunk consecutive tiles, the runs dealt round-robin to the CTAs.  chunk = 1 keeps all
    // CTAs on one front through the tensor (best L2/DRAM locality); with fused statistics longer runs keep a CTA
    // inside one image so that its running sums are flushed rarely.
    // PAIR: the work items are (cout tile, PAIR of pixel tiles); a cluster walks them, CTA rank r takes pixel tile 2*pair + r.
    const uint32_t rank = PAIR ? cluster_cta_rank() : 0u;
    const int walkers = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
    const int walker = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
    const int run_stride = walkers * P.chunk;
    auto tile_at = [&](int i) { return (i / P.chunk) * run_stride + walker * P.chunk + i % P.chunk; };
    struct Tile { int w0, h0, n0, c_out0; };
    // cout tile fastest: the CTAs that share an input box run at the same time (L2 reuse)
    auto decode = [&](int t) {
        Tile T;
        const int ct = t % P.cout_tiles;  t /= P.cout_tiles;
        if (PAIR) t = 2 * t + (int)rank;                       // may be one past the last pixel tile: every access is then out of bounds
        const int tw_i = t % P.n_tiles_w; t /= P.n_tiles_w;
        const int th_i = t % P.n_tiles_h; t /= P.n_tiles_h;
        T.w0 = tw_i * P.tw; T.h0 = th_i * P.cta_h; T.n0 = t * P.cta_n; T.c_out0 = ct * BN;
        return T;
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
        // PAIR: the even CTA's full barrier collects its own expect_tx arrival, the odd CTA's plain arrival and the bytes of
        // both CTAs' loads; its tempty barrier collects the epilogue threads of both CTAs
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), PAIR ? 2 : 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), PAIR ? 2 * kEpiThreads : kEpiThreads); }
        mbar_init(wres_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(smem_u32((const void*)tmem_slot)), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(smem_u32((const void*)tmem_slot)), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all(); else __syncthreads();        // barriers of both CTAs exist before anybody signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl::wait();             // from here on: activations in, outputs and statistics out

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer: runs ahead of the MMA by up to STAGES k-blocks, across tile boundaries ----
            uint32_t it = 0;                                     // global k-block counter -> stage / phase
            if (WRES) {                                          // the nine weight tiles, once (Cin == 64: K index = tap * 64)
                mbar_expect_tx(wres_bar, kWRes);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) tma_load_2d(wres + tap * kBBytes, &map_w, wres_bar, tap * BK, 0);
            }
            for (int ti = 0, tile = tile_at(0); tile < num_tiles; tile = tile_at(++ti)) {
                const Tile T = decode(tile);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t phase = (it / STAGES) & 1u;
                    mbar_wait(empty_bar(s), phase ^ 1u);
                    if (kBox) {                                  // one box: tile + halo in both directions
                        mbar_expect_tx(full_bar(s), (MT * 16 + 2) * 10 * 128);
                        tma_load_4d(base + s * kStageBytes, &map_x, full_bar(s), 0, T.w0 - 1, T.h0 - 1, T.n0);
                        continue;
                    }
                    if (HALO) {
                        const int cc = kb / 3, sx = kb - cc * 3;
                        const uint32_t a_dst = base + s * kStageBytes, b_dst = a_dst + kAHalo;
                        mbar_expect_tx(full_bar(s), kStageBytes);
                        tma_load_4d(a_dst, &map_x, full_bar(s), cc * BK, T.w0 + sx - 1, T.h0 - 1, T.n0);
                        if (!WRES) {
#pragma unroll
                            for (int r = 0; r < 3; ++r)
                                tma_load_2d(b_dst + r * kBBytes, &map_w, full_bar(s), ((r * 3 + sx) * P.cin_chunks + cc) * BK, T.c_out0);
                        }
                        continue;
                    }
                    const int tap = kb / P.cin_chunks, cc = kb - tap * P.cin_chunks;
                    const int r = tap / P.S, sx = tap - r * P.S;
                    const uint32_t a_dst = base + s * kStageBytes, b_dst = a_dst + MT * kABytes;
                    if (PAIR) {
                        if (rank == 0) mbar_expect_tx(full_bar(s), 2 * kStageBytes);
                        tma_load_4d_pair(a_dst, &map_x, full_bar(s), cc * BK, T.w0 * P.stride + sx - P.pad_w, T.h0 * P.stride + r - P.pad_h, T.n0);
                        tma_load_2d_pair(b_dst, &map_w, full_bar(s), kb * BK, T.c_out0 + (int)rank * (BN / 2));
                        if (rank != 0) mbar_arrive_even_cta(full_bar(s));
                        continue;
                    }
                    mbar_expect_tx(full_bar(s), kStageBytes);
#pragma unroll
                    for (int j = 0; j < MT; ++j)
                        tma_load_4d(a_dst + j * kABytes, &map_x, full_bar(s), cc * BK, T.w0 * P.stride + sx - P.pad_w,
                                    (T.h0 + j * P.sub_h) * P.stride + r - P.pad_h, T.n0 + j * P.sub_n);
                    tma_load_2d(b_dst, &map_w, full_bar(s), kb * BK, T.c_out0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ---- MMA issuer (PAIR: the even CTA only) ----
            uint32_t it = 0, local = 0;
            if (WRES) mbar_wait(wres_bar, 0);
            for (int ti = 0, tile = tile_at(0); tile < num_tiles; tile = tile_at(++ti), ++local) {
                const int as = local & 1;
                mbar_wait(tempty_bar(as), ((local >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)as * kAccCols;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t phase = (it / STAGES) & 1u;
                    mbar_wait(full_bar(s), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = base + s * kStageBytes;
                    if (kBox) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int r = tap / 3, sx = tap - r * 3;
                            const uint64_t bdesc = umma_desc_sw128(wres + (uint32_t)tap * kBBytes);
#pragma unroll
                            for (int k = 0; k < BK / UK; ++k)
#pragma unroll
                                for (int j = 0; j < MT; ++j) {
                                    const uint32_t a0 = a_addr + (uint32_t)((j * 16 + r) * 10 + sx) * 128u;
                                    const uint64_t adesc = umma_desc_sw128_sbo(a0, 1280u);          // image rows are 10 pixels = 1280 B apart
                                    umma_bf16(d_tmem + (uint32_t)(j * BN), adesc + (uint64_t)(k * UK * 2 / 16), bdesc + (uint64_t)(k * UK * 2 / 16),
                                              kIdesc, (uint32_t)((tap | k) != 0));
                                }
                        }
                        umma_commit(empty_bar(s));
                        continue;
                    }
                    if (HALO) {
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const uint64_t bdesc = WRES ? umma_desc_sw128(wres + (uint32_t)(r * 3 + (kb % 3)) * kBBytes)
                                                        : umma_desc_sw128(a_addr + kAHalo + r * kBBytes);
#pragma unroll
                            for (int k = 0; k < BK / UK; ++k)
#pragma unroll
                                for (int j = 0; j < MT; ++j)
                                    umma_bf16(d_tmem + (uint32_t)(j * BN), umma_desc_sw128(a_addr + (uint32_t)(j * 16 + r) * 1024u) + (uint64_t)(k * UK * 2 / 16),
                                              bdesc + (uint64_t)(k * UK * 2 / 16), kIdesc, (uint32_t)((kb | r | k) != 0));
                        }
                        umma_commit(empty_bar(s));
                        continue;
                    }
                    const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + MT * kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)
#pragma unroll
                        for (int j = 0; j < MT; ++j) {
                            if (PAIR)
                                umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * UK * 2 / 16), bdesc + (uint64_t)(k * UK * 2 / 16), kIdesc,
                                               (uint32_t)((kb | k) != 0));
                            else
                                umma_bf16(d_tmem + (uint32_t)(j * BN), adesc + (uint64_t)(j * (kABytes / 16) + k * UK * 2 / 16),
                                          bdesc + (uint64_t)(k * UK * 2 / 16), kIdesc, (uint32_t)((kb | k) != 0));
                        }
                    if (PAIR) umma_commit_pair(empty_bar(s)); else umma_commit(empty_bar(s));   // frees the stage (in both CTAs)
                }
                if (PAIR) umma_commit_pair(tfull_bar(as)); else umma_commit(tfull_bar(as));     // accumulator complete
            }
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 ----
        // Two warps per lane quarter: warp 2+q takes the first 32 columns of every 64-channel block, warp 6+q the second.
        const int q = warp & 3;
        const int eh = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const bool issuer = (warp == 2 && lane == 0);
        uint32_t blk = 0;                                        // staging blocks issued so far (buffer = blk & 1)
        uint32_t local = 0;
        // running InstanceNorm sums of this thread: [64-channel block][sum, sumsq of channel 2*lane, of 2*lane + 1]
        float st_acc[BN / 64][4];
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) st_acc[i][0] = st_acc[i][1] = st_acc[i][2] = st_acc[i][3] = 0.f;
        int st_key = -1;                                         // image * cout_tiles + cout tile the sums belong to
        // Flushing the running sums.  Every epilogue thread holds partial sums of the SAME BN channels (8 warps = 8 row groups),
        // and up to ~30 CTAs work on the same image: flushed thread by thread that was 4 096 fp64 atomics per CTA and flush on
        // 2 * BN addresses -- on small maps (a tile or two per CTA) slower than the statistics pass it replaces.  When a tile
        // never spans images (P.tn == 1: the key is CTA-uniform) the eight partials are first added in shared memory (fp32
        // atomics, 128 channels = 1 KB at a time), then ONE fp64 atomic per channel and statistic leaves the CTA.
        const bool st_uni = P.tn == 1;
        float* const st_red = reinterpret_cast<float*>(base_ptr + STAGES * kStageBytes + 2 * kOutBlk + 8 * (2 * STAGES + 6));   // [128][2]
        const int et = (int)threadIdx.x - 64;                    // 0 .. 255 among the epilogue threads
        auto stats_flush = [&]() {
            if (st_key >= 0) {
                const int n = st_key / P.cout_tiles, ct = st_key - n * P.cout_tiles;
                if (st_uni) {
#pragma unroll
                    for (int i0 = 0; i0 < BN / 64; i0 += 2) {
                        st_red[et] = 0.f;
                        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
                        for (int i = i0; i < i0 + 2 && i < BN / 64; ++i) {
                            float* d = st_red + ((i - i0) * 64 + 2 * lane) * 2;
                            atomicAdd(d, st_acc[i][0]); atomicAdd(d + 1, st_acc[i][1]); atomicAdd(d + 2, st_acc[i][2]); atomicAdd(d + 3, st_acc[i][3]);
                            st_acc[i][0] = st_acc[i][1] = st_acc[i][2] = st_acc[i][3] = 0.f;
                        }
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                        const int nch = (BN / 64 - i0 >= 2 ? 2 : 1) * 64;
                        if (et < 2 * nch && n < P.N) atomicAdd(P.stats + ((size_t)n * P.Cout + ct * BN + i0 * 64) * 2 + et, (double)st_red[et]);
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                    }
                    return;
                }
#pragma unroll
                for (int i = 0; i < BN / 64; ++i) {
                    double* dst = P.stats + ((size_t)n * P.Cout + ct * BN + i * 64 + 2 * lane) * 2;
                    atomicAdd(dst, (double)st_acc[i][0]); atomicAdd(dst + 1, (double)st_acc[i][1]);
                    atomicAdd(dst + 2, (double)st_acc[i][2]); atomicAdd(dst + 3, (double)st_acc[i][3]);
                    st_acc[i][0] = st_acc[i][1] = st_acc[i][2] = st_acc[i][3] = 0.f;
                }
            }
        };
        for (int ti = 0, tile = tile_at(0); tile < num_tiles; tile = tile_at(++ti), ++local) {
            const Tile T = decode(tile);
            const int as = local & 1;
            mbar_wait(tfull_bar(as), (local >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * kAccCols;
#pragma unroll 1
            for (int cb = 0; cb < MT * BN / 64; ++cb, ++blk) {         // cb walks sub-tile j = cb / (BN/64), then channels
                const int j = cb / (BN / 64), cblk = cb - j * (BN / 64);
                if (P.stats != nullptr && st_uni) {              // CTA-uniform: every epilogue thread takes part in the flush
                    const int key = (T.n0 + j * P.sub_n) * P.cout_tiles + (T.c_out0 / BN);
                    if (key != st_key) { stats_flush(); st_key = key; }
                }
                // the TMA store issued two blocks ago has finished reading this staging buffer
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                uint8_t* const srow = out_ptr + (blk & 1) * kOutBlk + row * 128;
                {
                    const int half = eh;
                    uint32_t v[32];
                    tmem_ld32(t_row + (uint32_t)(cb * 64 + half * 32), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (cb == MT * BN / 64 - 1) {
                        // last read of this accumulator: hand it back to the MMA warp before the stores
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        if (PAIR) mbar_arrive_even_cta(tempty_bar(as));
                        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(as)) : "memory");
                    }
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                    if (P.bias != nullptr) {
                        const float4* b4 = reinterpret_cast<const float4*>(P.bias + T.c_out0 + cblk * 64 + half * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 b = __ldg(b4 + i);
                            f[4 * i] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w;
                        }
                    }
                    if (P.slope != 1.0f) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.0f ? f[i] : f[i] * P.slope;
                    }
                    // staging block [128 rows][128 B]: 16-byte chunk j of row r sits at r*128 + ((j ^ (r & 7)) * 16),
                    // the layout a SWIZZLE_128B tensor map reads back
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int ch = half * 4 + i;
                        uint4 o;
                        o.x = pack_bf16(f[8 * i], f[8 * i + 1]);
                        o.y = pack_bf16(f[8 * i + 2], f[8 * i + 3]);
                        o.z = pack_bf16(f[8 * i + 4], f[8 * i + 5]);
                        o.w = pack_bf16(f[8 * i + 6], f[8 * i + 7]);
                        *reinterpret_cast<uint4*>(srow + ((ch ^ (row & 7)) << 4)) = o;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to TMA
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (issuer) {
                    tma_store_4d(&map_y, out_stage + (blk & 1) * kOutBlk, T.c_out0 + cblk * 64, T.w0, T.h0 + j * P.sub_h,
                                 T.n0 + j * P.sub_n);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (P.stats != nullptr) {
                    // InstanceNorm statistics of what was just stored (the bf16-rounded values, as a separate
                    // statistics pass over y would see them): thread = channel pair x one 32-row group of the block.
                    // One row of the staging block is 128 contiguous bytes -> a warp's 32 words never conflict.
                    // Sums stay in registers across blocks and tiles and go to memory only when the image or the
                    // channel tile changes (the CTA's tiles are consecutive, so that is rare).
                    const int r_mine = q * 32 + lane;                                  // lane rr describes row q*32 + rr
                    const int rows_per_img = P.tw * P.th;
                    const int in_img = r_mine % rows_per_img;
                    const int n_mine = T.n0 + j * P.sub_n + r_mine / rows_per_img;
                    const bool ok_mine = n_mine < P.N && (T.h0 + j * P.sub_h + in_img / P.tw) < P.Ho && (T.w0 + in_img % P.tw) < P.Wo;
                    const uint32_t valid = (__ballot_sync(0xffffffffu, ok_mine) >> (eh * 16)) & 0xffffu;   // this warp's 16 rows
                    const uint8_t* const grp = out_ptr + (blk & 1) * kOutBlk + (q * 32 + eh * 16) * 128 + (lane & 3) * 4;
                    const int chunk = lane >> 2;
                    if (valid == 0xffffu && P.tn == 1) {
                        // common case: the whole 32-row group lies inside one image -> 32 independent loads, then sums
                        uint32_t wv[16];
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr)
                            wv[rr] = *reinterpret_cast<const uint32_t*>(grp + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                        float p[4][4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) p[u][0] = p[u][1] = p[u][2] = p[u][3] = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            const float a = __uint_as_float(wv[rr] << 16), b = __uint_as_float(wv[rr] & 0xffff0000u);
                            float* pp = p[rr & 3];
                            pp[0] += a; pp[1] = fmaf(a, a, pp[1]); pp[2] += b; pp[3] = fmaf(b, b, pp[3]);
                        }
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)                              // static index: stays in registers
                            if (i == cblk) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) st_acc[i][e] += (p[0][e] + p[1][e]) + (p[2][e] + p[3][e]);
                            }
                    } else {
                        // ragged tile or a tile spanning several (tiny) images: row by row
#pragma unroll 1
                        for (int rr = 0; rr < 16; ++rr) {
                            if (!((valid >> rr) & 1u)) continue;                       // warp-uniform
                            if (!st_uni) {
                                const int n = __shfl_sync(0xffffffffu, n_mine, eh * 16 + rr);
                                const int key = n * P.cout_tiles + (T.c_out0 / BN);
                                if (key != st_key) { stats_flush(); st_key = key; }
                            }
                            const uint32_t word = *reinterpret_cast<const uint32_t*>(grp + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                            const float a = __uint_as_float(word << 16), b = __uint_as_float(word & 0xffff0000u);
#pragma unroll
                            for (int i = 0; i < BN / 64; ++i)
                                if (i == cblk) {
                                    st_acc[i][0] += a; st_acc[i][1] = fmaf(a, a, st_acc[i][1]);
                                    st_acc[i][2] += b; st_acc[i][3] = fmaf(b, b, st_acc[i][3]);
                                }
                        }
                    }
                }
            }
        }
        if (P.stats != nullptr) stats_flush();
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the reads
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all(); else __syncthreads();        // the peer may still be signalling this CTA's barriers / reading its smem
    if (warp == 2) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// `stride` > 1 (input map of a strided convolution): the box spans stride * bw x stride * bh input pixels and the TMA
// traversal stride (elementStrides) picks every stride-th of them, so shared memory receives the same bw x bh pixels --
// the strided im2col gather is done by the copy engine.
bool make_map_nhwc(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int bn, int bh, int bw, int stride = 1) {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
    const cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map_weights(CUtensorMap* m, const void* ptr, int cout, int k, int bn) {
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)cout};
    const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
    const cuuint32_t es[2] = {1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MT, int BN, int STAGES, bool PAIR = false, bool HALO = false, bool WRES = false>
cudaError_t launch(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const ConvParams& P,
                   long long ctas, cudaStream_t stream) {
    constexpr size_t stage = (WRES && STAGES == 2) ? (size_t)((((MT * 16 + 2) * 10 * 128) + 1023) / 1024 * 1024)
                           : WRES ? (size_t)(MT * 16 + 2) * 1024
                           : HALO ? (size_t)(MT * 16 + 2) * 1024 + 3 * (size_t)BN * BK * 2 : (size_t)(MT * kABytes + (PAIR ? BN / 2 : BN) * BK * 2);
    constexpr size_t smem = (WRES ? 9 * (size_t)BN * BK * 2 : 0) + (size_t)STAGES * stage + 2 * BM * 128 + 8 * (2 * STAGES + 6) + 1024 /* statistics flush scratch */ + 1024;
    static_assert(smem <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
    // The shared-memory opt-in and the SM count are PER DEVICE: cache them per device ordinal (per instantiation; the
    // attribute call is idempotent, so a race between host threads only repeats it).
    static bool attr_set[64] = {};
    static int sm_count[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cudaErrorInvalidDevice;
    const bool cached = dev >= 0 && dev < 64;
    if (!cached || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<MT, BN, STAGES, PAIR, HALO, WRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (cached) attr_set[dev] = true;
    }
    int sms = cached ? sm_count[dev] : 0;
    if (sms == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms = n;
        if (cached) sm_count[dev] = n;
    }
    // persistent: one CTA per SM walks the tile list with stride gridDim.x
    if (PAIR) {
        // clusters of two CTAs (one per SM of a TPC); `ctas` counts pixel tiles x cout tiles, a cluster takes two pixel tiles
        const long long items = (long long)P.cout_tiles * (((long long)P.n_tiles_w * P.n_tiles_h * P.n_tiles_n + 1) / 2);
        const long long clusters = items < sms / 2 ? items : sms / 2;
        return pdl::launch_cluster(conv_tc_kernel<MT, BN, STAGES, PAIR, HALO, WRES>, dim3((unsigned)(2 * clusters)), dim3(kThreads), smem, stream,
                                   2u, mx, mw, my, P);
    }
    const unsigned grid = (unsigned)(ctas < sms ? ctas : sms);
    return pdl::launch(conv_tc_kernel<MT, BN, STAGES, PAIR, HALO, WRES>, dim3(grid), dim3(kThreads), smem, stream, mx, mw, my, P);
}

// A/B switches for sweeps and parity tests (process-wide; atomics so that a concurrent launch reads a whole value, and every
// launch reads each switch ONCE -- a switch flipped mid-call cannot produce an inconsistent tile choice)
std::atomic<int> g_force_bn_sw{0};   // 0 = automatic; set through fots_b200_conv_set_tile for sweeps
std::atomic<int> g_halo_sw{-1};      // -1 = automatic, 0 = never, 1 = whenever the shape allows, 2 = same but always the three-copy form
                           // (fots_b200_conv_set_halo; sweeps / tests)

}  // namespace

extern "C" int fots_b200_conv_set_halo(int mode) {
    if (mode < -1 || mode > 2) return RROI_B200_ERR_INVALID_ARG;
    g_halo_sw.store(mode, std::memory_order_relaxed);
    return RROI_B200_OK;
}

extern "C" int fots_b200_conv_set_tile(int bn) {
    if (bn != 0 && bn != 64 && bn != 128 && bn != 256 && bn != 512) return RROI_B200_ERR_INVALID_ARG;   // 512 = 256 on a CTA pair
    g_force_bn_sw.store(bn, std::memory_order_relaxed);
    return RROI_B200_OK;
}

static int conv2d_impl(const void* x, const void* w, const float* bias, void* y, double* stats, int N, int H, int W,
                       int Cin, int Cout, int R, int S, int pad_h, int pad_w, float slope, cudaStream_t stream, int stride = 1) {
    if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || R <= 0 || S <= 0 || pad_h < 0 || pad_w < 0)
        return RROI_B200_ERR_INVALID_ARG;
    if (Cin % BK != 0 || Cout % 64 != 0 || R > 7 || S > 7 || (stride != 1 && stride != 2)) return RROI_B200_ERR_INVALID_ARG;
    if (H + 2 * pad_h < R || W + 2 * pad_w < S) return RROI_B200_ERR_INVALID_ARG;
    const int Ho = (H + 2 * pad_h - R) / stride + 1, Wo = (W + 2 * pad_w - S) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return RROI_B200_ERR_INVALID_ARG;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) return RROI_B200_ERR_INVALID_ARG;
    if (bias && ((uintptr_t)bias & 15)) return RROI_B200_ERR_INVALID_ARG;
    if (!encode_fn()) return RROI_B200_ERR_CUDA;

    // CTA pairs (cta_group::2) for 256-wide cout tiles: forced with tile = 512, automatic when there are enough pixel tiles
    // to keep all 74 clusters busy for several rounds (measured: conv8/9 1351 -> 1425 TF/s, conv7 1228 -> 1326; the
    // 256-tile conv10_s is better off with single CTAs)
    const int g_force_bn = g_force_bn_sw.load(std::memory_order_relaxed), g_halo = g_halo_sw.load(std::memory_order_relaxed);
    const long long px_tiles = ((long long)N * Ho * Wo + BM - 1) / BM;
    bool pair = Cout % 256 == 0 && stats == nullptr && (g_force_bn == 512 || (g_force_bn == 0 && px_tiles >= 4 * 148));
    int bn = pair ? 256 : (g_force_bn && g_force_bn != 512) ? g_force_bn : (Cout % 256 == 0 ? 256 : Cout % 128 == 0 ? 128 : 64);
    if (Cout % bn != 0) bn = 64;
    // BN < 256: two pixel sub-tiles per CTA share the weight tile (same bytes per MMA cycle as 128 x 256, and enough
    // MMAs per k-block to hide the single-thread issue path)
    const int mt = bn == 256 ? 1 : 2;

    // Halo reuse (three taps per A load; see the kernel): 3x3, stride 1, pad 1, 64- or 128-wide cout tiles.  Automatic for
    // maps of at least 32 x 16 pixels (the 8 x 32 CTA tile is then mostly inside the image): the backbone's stages, not
    // the recogniser's 8- and 4-row RoI tensors.
    const bool halo_ok = R == 3 && S == 3 && pad_h == 1 && pad_w == 1 && stride == 1 && stats == nullptr && !pair && (bn == 64 || bn == 128);
    const bool halo = halo_ok && (g_halo >= 1 || (g_halo == -1 && Ho >= 32 && Wo >= 16));
    // 64 -> 64 channels: the weights stay resident and ONE 10-pixel-wide box per tile carries all nine taps
    const bool halo_box = halo && g_halo != 2 && bn == 64 && Cin == 64 && Cout == 64;

    // sub-tile box: tw*th*tn = 128 pixels; the CTA stacks mt of them along h (tn == 1) or n.  Fewest CTA tiles wins,
    // wider boxes on ties (longer contiguous runs per TMA row).
    int best_tw = 0, best_th = 0, best_tn = 0;
    long long best = -1;
    for (int tw = 128; tw >= 8; tw >>= 1)
        for (int th = 128 / tw; th >= 1; th >>= 1) {
            if (tw * stride > 256 || th * stride > 256) continue;          // TMA box edge limit
            const int tn = 128 / (tw * th);
            const int ch = tn > 1 ? th : th * mt, cn = tn > 1 ? tn * mt : tn;
            const long long tiles = (long long)((Wo + tw - 1) / tw) * ((Ho + ch - 1) / ch) * ((N + cn - 1) / cn);
            if (best < 0 || tiles < best) { best = tiles; best_tw = tw; best_th = th; best_tn = tn; }
        }

    if (halo) { best_tw = 8; best_th = 16; best_tn = 1; }
    ConvParams P;
    P.tw = best_tw; P.th = best_th; P.tn = best_tn;
    const bool stack_n = best_tn > 1;
    P.cta_h = stack_n ? P.th : P.th * mt;  P.cta_n = stack_n ? P.tn * mt : P.tn;
    P.sub_h = stack_n ? 0 : P.th;          P.sub_n = stack_n ? P.tn : 0;
    P.n_tiles_w = (Wo + P.tw - 1) / P.tw; P.n_tiles_h = (Ho + P.cta_h - 1) / P.cta_h; P.n_tiles_n = (N + P.cta_n - 1) / P.cta_n;
    P.cin_chunks = Cin / BK; P.S = S; P.taps = R * S; P.pad_h = pad_h; P.pad_w = pad_w; P.stride = stride;
    P.cout_tiles = Cout / bn; P.slope = slope; P.bias = bias;
    P.stats = stats; P.N = N; P.Ho = Ho; P.Wo = Wo; P.Cout = Cout;
    P.chunk = 1;
    if (stats) {
        const cudaError_t em = pdl::zero_f64(stats, (size_t)N * Cout * 2, stream);
        if (em != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    }
    const long long ctas = (long long)P.n_tiles_w * P.n_tiles_h * P.n_tiles_n * P.cout_tiles;
    if (ctas > 0x7fffffffLL) return RROI_B200_ERR_TOO_LARGE;
    if (stats) {                           // runs of up to 8 tiles, but keep at least two runs per CTA for balance
        const long long c = ctas / (2 * 148);
        P.chunk = (int)(c < 1 ? 1 : c > 8 ? 8 : c);
    }

    CUtensorMap mx, mw, my;
    if (halo) {
        if (!make_map_nhwc(&mx, x, N, H, W, Cin, 1, mt * 16 + 2, halo_box ? 10 : 8, 1)) return RROI_B200_ERR_INVALID_ARG;      // tile + halo rows
    } else if (!make_map_nhwc(&mx, x, N, H, W, Cin, P.tn, P.th, P.tw, stride)) return RROI_B200_ERR_INVALID_ARG;
    if (!make_map_weights(&mw, w, Cout, R * S * Cin, pair ? bn / 2 : bn)) return RROI_B200_ERR_INVALID_ARG;
    if (!make_map_nhwc(&my, y, N, Ho, Wo, Cout, P.tn, P.th, P.tw)) return RROI_B200_ERR_INVALID_ARG;

    cudaError_t e;
    if (halo_box) e = launch<2, 64, 2, false, true, true>(mx, mw, my, P, ctas, stream);                   // one halo box, weights resident
    else if (halo && bn == 64 && Cin == 64 && Cout == 64) e = launch<2, 64, 3, false, true, true>(mx, mw, my, P, ctas, stream);   // weights resident
    else if (halo && bn == 64) e = launch<2, 64, 3, false, true>(mx, mw, my, P, ctas, stream);
    else if (halo) e = launch<2, 128, 2, false, true>(mx, mw, my, P, ctas, stream);
    else if (bn == 64) e = launch<2, 64, 4>(mx, mw, my, P, ctas, stream);
    else if (bn == 128) e = launch<2, 128, 4>(mx, mw, my, P, ctas, stream);
    else if (pair) e = launch<1, 256, 6, true>(mx, mw, my, P, ctas, stream);
    else e = launch<1, 256, 4>(mx, mw, my, P, ctas, stream);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

extern "C" int fots_b200_conv2d_nhwc_bf16(const void* x, const void* w, const float* bias, void* y, int N, int H, int W,
                                          int Cin, int Cout, int R, int S, int pad_h, int pad_w, float slope,
                                          cudaStream_t stream) {
    return conv2d_impl(x, w, bias, y, nullptr, N, H, W, Cin, Cout, R, S, pad_h, pad_w, slope, stream);
}

extern "C" int fots_b200_conv2d_stats_nhwc_bf16(const void* x, const void* w, const float* bias, void* y, double* stats,
                                                int N, int H, int W, int Cin, int Cout, int R, int S, int pad_h, int pad_w,
                                                cudaStream_t stream) {
    if (!stats) return RROI_B200_ERR_INVALID_ARG;
    return conv2d_impl(x, w, bias, y, stats, N, H, W, Cin, Cout, R, S, pad_h, pad_w, 1.0f, stream);
}

// Same convolution with a spatial stride of 1 or 2 (tools/models.py:257-264 layer0_1's second convolution, :283-296 the
// first block of every residual stage and its 1x1 down-sampling branch): the TMA tensor map's traversal stride gathers
// every second input pixel, everything else is the kernel above.
extern "C" int fots_b200_conv2d_strided_nhwc_bf16(const void* x, const void* w, const float* bias, void* y, int N, int H, int W,
                                                  int Cin, int Cout, int R, int S, int pad_h, int pad_w, int stride,
                                                  float slope, cudaStream_t stream) {
    return conv2d_impl(x, w, bias, y, nullptr, N, H, W, Cin, Cout, R, S, pad_h, pad_w, slope, stream, stride);
}
