// K1: float32 GEMM on the 5th-generation tensor cores — tcgen05.mma kind::tf32 with the accumulator
// in TMEM, operands staged in shared memory by TMA, and a 3xTF32 operand split so that the result
// stays float32-class (north_star: <= 1e-5 relative to OpenBLAS sgemm).
//
//   C(m x n) = op(A) op(B) [+ bias[i]]        Mul_M_M / Mul_M_M_Add_V_MCols (Backend.fs:119,121)
//
// The operand split.  kind::tf32 reads fp32 words from shared memory and IGNORES the low 13 mantissa
// bits, so the "hi" operand needs no plane of its own: the tensor core is pointed at the source buffer
// and sees hi = trunc_tf32(x).  Only lo = rna_tf32(x - trunc_tf32(x)) (x - hi is exact in fp32) has to
// be produced, in the SAME layout as the source, and no transposes are needed: an operand whose K index
// is not contiguous (B of an NN product, A of a TN product — every weight adjoint GEMM,
// AD.Float32.fs:4127-4143) is fed to the tensor core as an MN-major tile.  lo comes from
//   * a pre-pass (split_pair_kernel, HBM-bound, one launch for both operands: 4 bytes read + 4 written
//     per element) — the large products, or
//   * converter warps inside the kernel (single-CTA 128 x 128 / 128 x 64 tiles: the small products,
//     where a pre-pass launch costs more than it saves).
// Operands TMA cannot address directly (window not 16-byte aligned, contiguous extent not a multiple of
// 4) take the general route: hi and lo are both written by the pre-pass as zero-padded K-major planes
// [rows x Kp], transposed through shared memory on the way if needed.
//
// The kernel (gemm_tf32x3_kernel<BN, PAIR, WIDE>): persistent, warp-specialised.
//   warp 0     TMA producer: per 32-wide k-block the tiles of A and B into a ring of shared-memory stages
//              (K-major operand: one 32 x rows box per plane; MN-major operand: one 32(MN) x 32(K) box
//              per 32-wide MN block), completion on a `full` mbarrier
//   warp 1     allocates TMEM and issues tcgen05.mma from one thread: per 8-wide k-slice
//              D += A_lo*B_hi; D += A_hi*B_lo; D += A_hi*B_hi  (fp32 accumulator in TMEM);
//              tcgen05.commit releases the smem stage (`empty`) and, after the last k-block of a chunk,
//              publishes the accumulator (`tmem_full`)
//   8 warps    epilogue: tcgen05.ld the accumulator (each warp owns the 32 TMEM lanes of its quadrant and
//              half of the columns), add the chunk into registers, and after the last chunk add the bias
//              and store float4s to C.  Two accumulator stages (2 x BN TMEM columns) let the drain of
//              chunk c overlap the MMAs of chunk c+1.
//   (+ 4 / 6)  converter warps in the variants that split in the kernel.
// PAIR: the two CTAs of a cluster (the two SMs of a TPC) run ONE tcgen05.mma.cta_group::2 on a 256 x BN
// tile — see Cfg below; this is the variant the planner picks for every large product.
// The a_lo*b_lo term (<= 2^-22 relative) is dropped: that is the "3x" in 3xTF32.
//
// Accuracy of long reductions.  The tensor core's fp32 accumulator TRUNCATES on every accumulate
// (measured on B200: same-sign data, K = 8192, one TMEM accumulator => 9.1e-5 relative error,
// = 3 MMAs x K/16 x 2^-24).  The K loop is therefore cut into chunks of kChunkKb k-blocks (256
// elements): each chunk accumulates from zero in TMEM (truncation bias <= 2.8e-6 of the chunk), and
// the epilogue warps add the chunk results in registers with IEEE round-to-nearest fp32 adds.
//
// Schedule.  The work of a GEMM is the list of (tile, K-chunk) units in tile-major order, shared out
// over one worker (CTA, or CTA pair) per SM (pair of SMs):
//   tiles >= workers : whole-tile waves, plus stream-K over the last, partial wave when K is deep;
//   tiles <  workers : aligned split-K — every tile is cut into equal K ranges.
// A run of units of one tile inside one worker is a SEGMENT.  A segment that covers its whole tile is
// stored to C directly.  Partial segments are combined in K order with IEEE adds, whichever path they
// take, so the result does not depend on timing (deterministic):
//   * aligned split-K whose K ranges fit one thread-block cluster: the partial accumulators cross
//     distributed shared memory and every CTA finishes 1/nseg of the columns from registers;
//   * other aligned splits: plane-major workspace slots; the CTAs of a tile wait for each other and each
//     reduces 1/nseg of the planes;
//   * stream-K ranges: workspace slots + ticket; the CTA that arrives last adds the tile's segments.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "dsc_internal.cuh"
#include "dsc_ops.cuh"

namespace dsc {
namespace tf32x3 {

constexpr int BM = 128;
constexpr int BK = 32;  // floats: 128 bytes = one SWIZZLE_128B row
constexpr int kEpiWarps = 8;
constexpr int kBaseThreads = 64 + 32 * kEpiWarps;  // warp0 TMA, warp1 MMA, 8 epilogue warps (+ converter warps, see Cfg)
constexpr int kChunkKb = 8;                   // k-blocks per accumulation chunk (256 K-elements)
constexpr int kAPlaneBytes = BM * BK * 4;

// PAIR: two CTAs of a cluster (the two SMs of a TPC) run ONE tcgen05.mma.cta_group::2 on a 256 x BN
// tile: each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows) and the tensor cores
// of both SMs share the B halves, so a CTA moves (128 + BN/2) rows per k-block through shared memory
// instead of (128 + BN).  That matters because shared-memory bandwidth, not the tensor pipe, bounds
// the single-CTA kernel: per 32-wide k-block a 128 x 256 tile takes 96 KB of TMA writes plus
// 12 MMAs x 12 KB of operand reads = 240 KB at 128 B/cycle = 1920 cycles against 1536 cycles of MMA
// (measured: tensor pipe 78 % active).  The pair moves 64 + 96 = 160 KB = 1280 cycles: tensor-bound.
//
// WIDE (256 x 256 pair tiles whose operands are both directly addressable): the operand split moves
// into the kernel.  TMA fetches only the fp32 words (the hi tiles: half the L2 -> SM bytes, no lo
// planes in HBM, no pre-pass launch) into a 4-deep ring; six converter warps derive the lo tiles into a
// separate 3-deep ring (7 x 32 KB of shared memory), so that the extra hop TMA -> converters -> MMA
// is hidden: the fetch of k-block i+3 and the conversion of k-blocks i+1, i+2 overlap the MMAs of
// k-block i.  16 warps in four warpgroups — {TMA, MMA, converter, converter}, 2 x epilogue,
// 4 x converter — with the register file re-divided by setmaxnreg (epilogue threads hold 128
// accumulators: 200 registers; the other warpgroups run on 56).  Per k-block a CTA moves 32 KB of TMA
// writes + 64 KB of converter traffic + 96 KB of MMA operand reads = 192 KB = 1536 cycles of shared-memory
// bandwidth, below the ~2400 cycles the tensor pipe needs for the 12 TF32 MMAs of a 256 x 256 x 32 block.
template <int BN, bool PAIR, bool WIDE>
struct Cfg {
    static_assert(!WIDE || (PAIR && BN == 256), "the in-kernel split of the pair kernel is built for 256 x 256 tiles");
    static constexpr bool kWide = WIDE;
    static constexpr int kBRows = PAIR ? BN / 2 : BN;     // rows of the B tile staged by one CTA
    static constexpr int kBPlaneBytes = kBRows * BK * 4;
    // not WIDE: a stage holds [A_hi | A_lo | B_hi | B_lo];  WIDE: hi ring of [A_hi | B_hi] stages, then the lo ring of [A_lo | B_lo]
    static constexpr int kStageBytes = WIDE ? kAPlaneBytes + kBPlaneBytes : 2 * kAPlaneBytes + 2 * kBPlaneBytes;
    static constexpr int kStages = WIDE ? 4 : ((196608 / kStageBytes) > 8 ? 8 : (196608 / kStageBytes));
    static constexpr int kLoStages = WIDE ? 3 : kStages;
    static constexpr int kTileBytes = WIDE ? (kStages + kLoStages) * kStageBytes : kStages * kStageBytes;
    static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator stages; power of two
    static constexpr int kSmemBytes = kTileBytes + 1024 /*align*/ + 256 /*barriers, TMEM base, epilogue flag*/;
    // Converter warps (in-kernel operand split).  Single-CTA 128 x 256 tiles leave room for two smem
    // stages only and need 168 registers per epilogue thread: that shape keeps the pre-pass, and so do
    // the narrower pair tiles (measured: with converter warps and the extra TMA -> converter -> MMA hop
    // the 256 x 128 pair kernel loses 2-3 us on a 1024^3 product even when nothing is converted);
    // single-CTA 128 x 128 / 128 x 64 tiles convert inside their stages.
    static constexpr int kConvWarps = WIDE ? 6 : ((PAIR || BN >= 256) ? 0 : 4);
    static constexpr int kEpi0 = WIDE ? 4 : 2;   // first epilogue warp
    static constexpr int kThreads = WIDE ? 512 : kBaseThreads + 32 * kConvWarps;
};

// ---- PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
// 2-CTA forms: the load lands in this CTA's shared memory, its bytes are counted on the LEADER's
// (even CTA's) barrier: clearing bit 24 of a shared::cta address names the same offset in the even CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar, uint16_t mask) {   // arrives on `bar` in BOTH CTAs of the pair (mask = their cluster ranks)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope: the arrivals come from both CTAs
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_leader_release(uint32_t bar, uint32_t leader) {   // release at cluster scope on the barrier of cluster CTA `leader`
    asm volatile(
        "{\n\t"
        ".reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [rem];\n\t"
        "}" ::"r"(bar), "r"(leader)
        : "memory");
}
// 16 bytes into the shared memory of cluster CTA `rank`, at the offset `addr` has in this CTA (distributed shared memory)
__device__ __forceinline__ void st_dsmem_f4(uint32_t addr, uint32_t rank, float4 v) {
    asm volatile(
        "{\n\t"
        ".reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
        "st.shared::cluster.v4.f32 [rem], {%2, %3, %4, %5};\n\t"
        "}" ::"r"(addr), "r"(rank), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar, uint32_t leader) {    // arrive on the barrier at the same offset in cluster CTA `leader`
    asm volatile(
        "{\n\t"
        ".reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t"
        "}" ::"r"(bar), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives columns [c, c+32) of TMEM lane base+i
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major SWIZZLE_128B tile (rows of 128 B, 8-row groups of
// 1024 B): start address >> 4 in bits [0,14), LBO (unused for swizzled K-major) = 1 in [16,30),
// SBO = 1024 B >> 4 in [32,46), descriptor version 1 in [46,48), layout SWIZZLE_128B (=2) in [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// MN-major tile of a 32-bit operand: [MN/32 blocks][32 k-rows][32 floats]; a k-row of a block is
// 128 B (32 MN-contiguous elements).  For 32-bit MN-major operands the tensor core accepts one
// layout only, SWIZZLE_128B_BASE32B (descriptor layout type 1; TMA: SWIZZLE_128B_ATOM_32B): 32-byte
// chunks of a 128-byte row are XOR-ed with (row % 4), so the swizzle atom is 4 k-rows = 512 B.
// Canonical form ((4,8,m),(4,k)):((1,4,LBO),(32,SBO)) in elements: LBO = distance between 32-wide
// MN blocks (32 k-rows x 128 B = 4096 B), SBO = distance between 4-row k groups (512 B).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(4096 >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

// Instruction descriptor, kind::tf32: D = F32 (bits 4-5 = 1), A/B = TF32 (bits 7-9, 10-12 = 2),
// A / B major-ness in bits 15 / 16 (0 = K-major, 1 = MN-major), N >> 3 in bits [17,23), M >> 4 in
// bits [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Params {
    int m, n, num_kb;          // problem rows/cols, number of 32-wide k-blocks
    int tiles_m, tiles_n;
    int cpt;                   // K chunks per tile
    // schedule: CTA b first runs dp_waves whole tiles (tile = wave * grid + b), then its share of the
    // stream-K units of the remaining tiles [dp_tiles, tiles): base + (b < rem) units from
    // b*base + min(b, rem), unit u = (dp_tiles + u / cpt, u % cpt)
    int dp_waves, dp_tiles;
    int dp_extra;              // CTAs b < dp_extra run one more whole tile (tile = dp_waves * grid + b): plain last wave
    int perm_tiles, perm_s;    // aligned split-K: CTA b acts as logical CTA (b % perm_tiles) * perm_s + b / perm_tiles, so that
                               // neighbouring CTAs work on neighbouring tiles of the same K range (0: identity)
    int base, rem;
    float* C;
    const float* bias;
    float* slots;              // [grid][2][BM x BN] partial segment sums
    unsigned int* tickets;     // one arrival counter per tile (zero outside a launch)
    int vec4;                  // output rows are 16-byte aligned and n % 4 == 0
    int a_mn, b_mn;            // operand tiles are MN-major (source read in place, K index strided)
    int a_fused, b_fused;      // the operand's lo tile is computed in shared memory by the converter warps (no lo plane in HBM)
    int kcluster;              // > 1: aligned split-K whose K ranges of a tile are the CTAs (pairs) of ONE cluster; the partial
                               // tiles are exchanged through distributed shared memory instead of workspace slots
    unsigned long long* stamps; // null, or 8 slots for %globaltimer at the phase boundaries as CTA 0 sees them (dsc_gemm_debug_stamps)
    // Epilogue (GemmEpilogue, dsc_internal.cuh): what the calls that FOLLOW Mul_M_M in the unchanged sequence do with the
    // product, applied to the finished tile in registers — same roundings as the separate kernels (dsc_ops.cuh)
    int ep_mode, ep_op;
    const float* ep_vec;       // GEMM_EP_BIAS_ACT: v[j]
    const float* ep_aux;       // GEMM_EP_CHAIN: p[i, j]
    float* ep_out2;            // GEMM_EP_BIAS_ACT with an activation: h
    float* ep_lo;              // operand-split companion of the last output, or null
};
__device__ __forceinline__ void stamp(const Params& p, int slot) {
    if (p.stamps != nullptr && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.stamps[slot] = t;
    }
}

__device__ __forceinline__ int unit_begin(const Params& p, int b) { return b * p.base + min(b, p.rem); }
__device__ __forceinline__ int unit_owner(const Params& p, int u) {
    const int big = p.rem * (p.base + 1);
    return u < big ? u / (p.base + 1) : p.rem + (u - big) / p.base;
}
// i-th unit of CTA b -> (tile, chunk)
__device__ __forceinline__ void unit_at(const Params& p, int i, int dp_units, int sk_begin, int nb, int bid, int& tile, int& chunk) {
    if (i < dp_units) {
        tile = (i / p.cpt) * nb + bid;  // whole-tile waves use the physical CTA (pair) index
        chunk = i % p.cpt;
    } else {
        const int u = sk_begin + (i - dp_units);
        tile = p.dp_tiles + u / p.cpt;
        chunk = u % p.cpt;
    }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }

// ---- the finished tile leaves the registers ---------------------------------------------------------
// v = four consecutive elements of row `row` (< m) starting at column `col`, bias[i] of Mul_M_M_Add_V_MCols already added.
//   GEMM_EP_NONE    : C = v
//   GEMM_EP_BIAS_ACT: z = v + vec[col..] -> C;  h = f_op(z) -> out2 (op < 0: none)     Add_M_M(., rows(vec)) -> Map_F_M
//   GEMM_EP_CHAIN   : y = chain_op(v, aux[row, col..]) -> C                             the adjoint chains
// and, when asked for, lo = dsc_lo_of(last output): the operand-split companion for the GEMM that reads it next.
// Activations fused here: Tanh, Sigmoid, ReL (the host runs any other op-code as a second pass over C).
// The epilogue arithmetic is kept OUT of line: the finished tile leaves the registers as fully unrolled straight-line code
// (the accumulators are register-indexed), and inlining tanhf / coshf per element there grows the kernel past the
// instruction cache (measured: 2.5x slower products).  One call handles 16 consecutive columns of a row: its loads are issued
// together, then the math, then the stores.
struct EpArgs {           // what the out-of-line epilogue needs of Params, by value (registers)
    float* C;
    const float* vec;
    const float* aux;
    float* out2;
    float* lo;
    int n, vec4, mode, op;
};
__device__ __forceinline__ EpArgs ep_args(const Params& p) { return EpArgs{p.C, p.ep_vec, p.ep_aux, p.ep_out2, p.ep_lo, p.n, p.vec4, p.ep_mode, p.ep_op}; }

template <int OP, int N>
__device__ __forceinline__ void ep_bias_act(const EpArgs& p, size_t base, int col, float (&v)[N], bool full) {
    float b[N];
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
        if (full) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.vec + col) + q);
            b[4 * q] = t.x; b[4 * q + 1] = t.y; b[4 * q + 2] = t.z; b[4 * q + 3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) b[4 * q + j] = (col + 4 * q + j < p.n) ? __ldg(p.vec + col + 4 * q + j) : 0.f;
        }
    }
    float h[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        v[j] = v[j] + b[j];
        h[j] = OP >= 0 ? unary_op<(OP >= 0 ? OP : DSC_OP_TANH)>(v[j]) : v[j];
    }
    if (full) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            reinterpret_cast<float4*>(p.C + base)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            if (OP >= 0) reinterpret_cast<float4*>(p.out2 + base)[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
            if (p.lo) reinterpret_cast<float4*>(p.lo + base)[q] = make_float4(dsc_lo_of(h[4 * q]), dsc_lo_of(h[4 * q + 1]), dsc_lo_of(h[4 * q + 2]), dsc_lo_of(h[4 * q + 3]));
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (col + j < p.n) {
                p.C[base + j] = v[j];
                if (OP >= 0) p.out2[base + j] = h[j];
                if (p.lo) p.lo[base + j] = dsc_lo_of(h[j]);
            }
    }
}
template <int KIND, int N>
__device__ __forceinline__ void ep_chain(const EpArgs& p, size_t base, int col, float (&v)[N], bool full) {
    float a[N];
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
        if (full) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.aux + base) + q);
            a[4 * q] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) a[4 * q + j] = (col + 4 * q + j < p.n) ? __ldg(p.aux + base + 4 * q + j) : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = chain_op<KIND>(v[j], a[j]);
    if (full) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            reinterpret_cast<float4*>(p.C + base)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            if (p.lo) reinterpret_cast<float4*>(p.lo + base)[q] = make_float4(dsc_lo_of(v[4 * q]), dsc_lo_of(v[4 * q + 1]), dsc_lo_of(v[4 * q + 2]), dsc_lo_of(v[4 * q + 3]));
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (col + j < p.n) {
                p.C[base + j] = v[j];
                if (p.lo) p.lo[base + j] = dsc_lo_of(v[j]);
            }
    }
}
template <int N>
__device__ __forceinline__ void ep_dispatch(const EpArgs& p, int row, int col, float (&v)[N]) {
    const size_t base = (size_t)row * p.n + col;
    const bool full = p.vec4 && col + N <= p.n;
    if (p.mode == GEMM_EP_BIAS_ACT) {
        switch (p.op) {
            case DSC_OP_TANH: ep_bias_act<DSC_OP_TANH, N>(p, base, col, v, full); break;
            case DSC_OP_SIGMOID: ep_bias_act<DSC_OP_SIGMOID, N>(p, base, col, v, full); break;
            case DSC_OP_REL: ep_bias_act<DSC_OP_REL, N>(p, base, col, v, full); break;
            default: ep_bias_act<-1, N>(p, base, col, v, full); break;
        }
    } else {
        switch (p.op) {
            case DSC_FUSED_TANH_BWD: ep_chain<DSC_FUSED_TANH_BWD, N>(p, base, col, v, full); break;
            case DSC_FUSED_SIGMOID_BWD: ep_chain<DSC_FUSED_SIGMOID_BWD, N>(p, base, col, v, full); break;
            default: ep_chain<DSC_FUSED_RELU_BWD, N>(p, base, col, v, full); break;
        }
    }
}
// 16 / 4 consecutive elements of row `row` (< m) from column `col`, bias[i] of Mul_M_M_Add_V_MCols already added, under a
// non-trivial epilogue
__device__ __noinline__ void store16_ep(EpArgs p, int row, int col, float4 q0, float4 q1, float4 q2, float4 q3) {
    if (col >= p.n) return;
    float v[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
    ep_dispatch<16>(p, row, col, v);
}
__device__ __noinline__ void store4_ep(EpArgs p, int row, int col, float4 q0) {
    if (col >= p.n) return;
    float v[4] = {q0.x, q0.y, q0.z, q0.w};
    ep_dispatch<4>(p, row, col, v);
}
// four consecutive elements: plain store, or the epilogue out of line
__device__ __forceinline__ void store_quad(const Params& p, int row, int col, float v0, float v1, float v2, float v3) {
    if (col >= p.n) return;
    if (p.ep_mode != GEMM_EP_NONE) {
        store4_ep(ep_args(p), row, col, make_float4(v0, v1, v2, v3));
        return;
    }
    float* dst = p.C + (size_t)row * p.n + col;
    if (p.vec4 && col + 4 <= p.n) *reinterpret_cast<float4*>(dst) = make_float4(v0, v1, v2, v3);
    else {
        dst[0] = v0;
        if (col + 1 < p.n) dst[1] = v1;
        if (col + 2 < p.n) dst[2] = v2;
        if (col + 3 < p.n) dst[3] = v3;
    }
}

// ---- in-kernel operand split ---------------------------------------------------------------------
// lo = rna_tf32(x - trunc_tf32(x)) for one fp32 word as the tensor core will see it: x - hi is exact;
// adding half a TF32 ulp to the bit pattern rounds the magnitude to nearest once the tensor core
// drops the low 13 bits.  Non-finite x (x - hi = NaN) contributes through hi only.
__device__ __forceinline__ float lo_word(float x) {
    const float res = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    return res == res ? __uint_as_float(__float_as_uint(res) + 0x1000u) : 0.f;
}
// lo plane = f(hi plane), element by element (the swizzle is the same for both, so layout is irrelevant).
// N4 float4 per plane, NT converter threads, 4 loads in flight per thread.
template <int N4, int NT>
__device__ __forceinline__ void convert_plane(const uint8_t* hi, uint8_t* lo, int ct) {
    constexpr int PER = (N4 + NT - 1) / NT;   // float4 per thread (the last round may be partial)
    constexpr int U = 4;
#pragma unroll
    for (int j = 0; j < PER; j += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = ct + (j + u) * NT;
            if (j + u < PER && ((j + u + 1) * NT <= N4 || idx < N4)) v[u] = *reinterpret_cast<const float4*>(hi + 16 * idx);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = ct + (j + u) * NT;
            if (j + u < PER && ((j + u + 1) * NT <= N4 || idx < N4))
                *reinterpret_cast<float4*>(lo + 16 * idx) = make_float4(lo_word(v[u].x), lo_word(v[u].y), lo_word(v[u].z), lo_word(v[u].w));
        }
    }
}

template <int BN, bool PAIR, bool WIDE>
__global__ void __launch_bounds__(Cfg<BN, PAIR, WIDE>::kThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                   const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, const Params p) {
    using cfg = Cfg<BN, PAIR, WIDE>;
    constexpr int TM = PAIR ? 2 * BM : BM;          // rows of an output tile
    // cluster = [K ranges of one tile (kcluster, optional)] x [the two CTAs of a pair (PAIR)]
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const uint32_t cta_rank = PAIR ? (crank & 1u) : 0u;   // 0 = leader of the pair (issues the MMAs)
    const uint32_t pair_leader = crank & ~1u;             // cluster rank of this CTA's pair leader
    const uint32_t kidx = PAIR ? (crank >> 1) : crank;    // which K range of the tile (kcluster mode)
    const int nb = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;    // workers (CTAs or CTA pairs)
    const int bid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + cfg::kTileBytes;
    // barrier slots (8 bytes each): full[kStages], empty[kStages], conv[kLoStages], lo_empty[kLoStages], tmem_full[2],
    // tmem_empty[2], then the TMEM base and the epilogue flag
    constexpr int kConv = cfg::kConvWarps;
    constexpr int kEpi0 = cfg::kEpi0;
    constexpr int kS = cfg::kStages, kL = cfg::kLoStages;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kS + s); };
    auto conv_bar = [&](int l) { return bars + 8u * (2 * kS + l); };
    auto loempty_bar = [&](int l) { return bars + 8u * (2 * kS + kL + l); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * kS + 2 * kL + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * kS + 2 * kL + 2 + a); };
    constexpr int kSlotOff = 8 * (2 * kS + 2 * kL + 4);
    static_assert(kSlotOff + 16 <= 256, "barrier block");
    const uint32_t tmem_slot = bars + kSlotOff;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + cfg::kTileBytes + kSlotOff);
    // tile addresses (byte offsets from smem_base): s = fetch stage, l = lo-ring stage (not WIDE: the lo tiles live in stage s)
    auto a_hi_off = [&](int s) { return (uint32_t)(s * cfg::kStageBytes); };
    auto b_hi_off = [&](int s) { return (uint32_t)(s * cfg::kStageBytes + (WIDE ? kAPlaneBytes : 2 * kAPlaneBytes)); };
    auto a_lo_off = [&](int s, int l) { return (uint32_t)(WIDE ? (kS + l) * cfg::kStageBytes : s * cfg::kStageBytes + kAPlaneBytes); };
    auto b_lo_off = [&](int s, int l) {
        return (uint32_t)(WIDE ? (kS + l) * cfg::kStageBytes + kAPlaneBytes : s * cfg::kStageBytes + 2 * kAPlaneBytes + cfg::kBPlaneBytes);
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) stamp(p, 0);   // kernel entry
    if (threadIdx.x == 32) {             // descriptor fetches overlap the barrier / TMEM prologue
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ahi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_alo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo) : "memory");
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kS; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int l = 0; l < kL; ++l) {
            mbar_init(conv_bar(l), kConv > 0 ? (PAIR ? 2 * kConv : kConv) : 1);  // one arrival per converter warp (pair: of both CTAs, on the leader's barrier)
            mbar_init(loempty_bar(l), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), PAIR ? 2 * kEpiWarps : kEpiWarps);  // one arrival per epilogue warp (pair: of both CTAs, on the leader's barrier)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)cfg::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)cfg::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (PAIR || p.kcluster > 1) cluster_sync_all();   // the peers run and their barriers are initialised before anything arrives there
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    if (threadIdx.x == 0) stamp(p, 1);   // barriers initialised, TMEM allocated

    const int lb = p.perm_tiles ? (bid % p.perm_tiles) * p.perm_s + bid / p.perm_tiles : bid;
    const int sk_begin = unit_begin(p, lb), sk_end = unit_begin(p, lb + 1);
    const int dp_units = (p.dp_waves + (bid < p.dp_extra ? 1 : 0)) * p.cpt;
    const int my_units = dp_units + (sk_end - sk_begin);
    const bool a_fused = WIDE || (kConv > 0 && p.a_fused), b_fused = WIDE || (kConv > 0 && p.b_fused);   // WIDE is launched for two fused operands only
    const uint32_t stage_tx = (uint32_t)((a_fused ? 1 : 2) * kAPlaneBytes + (b_fused ? 1 : 2) * cfg::kBPlaneBytes);

    // ===================== converter warps: lo tiles from hi tiles, in shared memory =====================
    auto run_converter = [&]() {
        constexpr int NT = 32 * (kConv > 0 ? kConv : 1);
        // converter thread index: the warps before the epilogue block (from warp 2), then the warps after it
        const int ct = warp < kEpi0 ? (int)threadIdx.x - 64 : (int)threadIdx.x - 32 * (kEpi0 + kEpiWarps) + 32 * (kEpi0 - 2);
        int s = 0, l = 0;
        uint32_t ph = 0, lph = 0;
        for (int i = 0; i < my_units; ++i) {
            int tile, chunk;
            unit_at(p, i, dp_units, sk_begin, nb, bid, tile, chunk);
            const int kb0 = chunk * kChunkKb;
            const int kb1 = min(p.num_kb, kb0 + kChunkKb);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full_bar(s), ph);                              // this CTA's fp32 tiles have landed
                if constexpr (WIDE) mbar_wait(loempty_bar(l), lph ^ 1u); // the MMAs that read this lo-ring stage have retired
                if (a_fused) convert_plane<kAPlaneBytes / 16, NT>(smem_gen + a_hi_off(s), smem_gen + a_lo_off(s, l), ct);
                if (b_fused) convert_plane<cfg::kBPlaneBytes / 16, NT>(smem_gen + b_hi_off(s), smem_gen + b_lo_off(s, l), ct);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_leader_release(conv_bar(l), pair_leader);
                    else mbar_arrive(conv_bar(l));
                }
                if (++s == kS) { s = 0; ph ^= 1u; }
                if (++l == kL) { l = 0; lph ^= 1u; }
            }
        }
    };

    // WIDE: the register file is re-divided per warpgroup; every warp of a warpgroup executes the SAME
    // setmaxnreg instruction (warpgroup 0 here, warpgroups 1-2 at the top of the epilogue branch,
    // warpgroup 3 at the top of the converter branch), and ptxas sizes each region accordingly.
    if constexpr (cfg::kWide) {
        if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    }
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < my_units; ++i) {
                int tile, chunk;
                unit_at(p, i, dp_units, sk_begin, nb, bid, tile, chunk);
                // this CTA's share of the tile's operands: its 128 rows of A, and (pair) its half of the B rows
                const int m0 = (tile / p.tiles_n) * TM + (int)cta_rank * BM, n0 = (tile % p.tiles_n) * BN + (int)cta_rank * cfg::kBRows;
                const int kb0 = chunk * kChunkKb;
                const int kb1 = min(p.num_kb, kb0 + kChunkKb);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t a_hi = smem_base + a_hi_off(s), a_lo = smem_base + a_lo_off(s, s);
                    const uint32_t b_hi = smem_base + b_hi_off(s), b_lo = smem_base + b_lo_off(s, s);   // lo tiles are fetched only when not fused (never WIDE)
                    // Pair without converters: the bytes of both CTAs are counted on the LEADER's barrier (the
                    // MMA issuer waits there).  With converters every CTA counts its own bytes on its own
                    // barrier: its converter warps wait for them, and the leader's MMA issuer waits for the
                    // converter warps of both CTAs.
                    constexpr bool kLeaderFull = PAIR && kConv == 0;
                    auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1) {
                        if constexpr (kLeaderFull) tma_load_2d_pair(dst, map, full_bar(s), c0, c1);
                        else tma_load_2d(dst, map, full_bar(s), c0, c1);
                    };
                    if (!kLeaderFull || cta_rank == 0) mbar_expect_tx(full_bar(s), kLeaderFull ? 2u * stage_tx : stage_tx);
                    if (!p.a_mn) {
                        load(a_hi, &map_ahi, kb * BK, m0);
                        if (!a_fused) load(a_lo, &map_alo, kb * BK, m0);
                    } else {
#pragma unroll
                        for (int i = 0; i < BM / 32; ++i) {
                            load(a_hi + i * 4096, &map_ahi, m0 + 32 * i, kb * BK);
                            if (!a_fused) load(a_lo + i * 4096, &map_alo, m0 + 32 * i, kb * BK);
                        }
                    }
                    if (!p.b_mn) {
                        load(b_hi, &map_bhi, kb * BK, n0);
                        if (!b_fused) load(b_lo, &map_blo, kb * BK, n0);
                    } else {
#pragma unroll
                        for (int i = 0; i < cfg::kBRows / 32; ++i) {
                            load(b_hi + i * 4096, &map_bhi, n0 + 32 * i, kb * BK);
                            if (!b_fused) load(b_lo + i * 4096, &map_blo, n0 + 32 * i, kb * BK);
                        }
                    }
                    if (++s == kS) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0 && cta_rank == 0) {
            const uint32_t idesc = make_idesc_tf32(TM, BN) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16);
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
                if constexpr (PAIR) umma_tf32_pair(d, a, b, idesc, acc);
                else umma_tf32(d, a, b, idesc, acc);
            };
            auto commit = [&](uint32_t bar) {
                if constexpr (PAIR) tcgen05_commit_pair(bar, (uint16_t)(3u << pair_leader));
                else tcgen05_commit(bar);
            };
            // per 8-wide k-slice the descriptor start address moves by 32 B inside the swizzle row
            // (K-major) or by one 1024-byte group of 8 k-rows (MN-major); units of 16 B
            const uint64_t a_adv = p.a_mn ? 64u : 2u, b_adv = p.b_mn ? 64u : 2u;
            int s = 0, l = 0, as = 0;
            uint32_t ph = 0, lph = 0, aph = 0;
            for (int i = 0; i < my_units; ++i) {
                int tile, chunk;
                unit_at(p, i, dp_units, sk_begin, nb, bid, tile, chunk);
                const int c0 = chunk * kChunkKb;
                const int c1 = min(p.num_kb, c0 + kChunkKb);
                mbar_wait(tempty_bar(as), aph ^ 1u);  // the epilogue has drained this accumulator stage
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = c0; kb < c1; ++kb) {
                    if constexpr (PAIR && kConv > 0) {
                        mbar_wait_cluster(conv_bar(l), lph);    // the converter warps of BOTH CTAs have seen their TMA bytes and written the lo tiles
                    } else {
                        mbar_wait(full_bar(s), ph);             // TMA bytes have landed
                        if (kConv > 0) mbar_wait(conv_bar(l), lph);  // ... and the converter warps have written the lo tiles
                    }
                    tcgen05_fence_after();
                    if (i == 0 && kb == c0) stamp(p, 2);   // first operand tiles have landed
                    const uint32_t a_hi = smem_base + a_hi_off(s), a_lo = smem_base + a_lo_off(s, l);
                    const uint32_t b_hi = smem_base + b_hi_off(s), b_lo = smem_base + b_lo_off(s, l);
                    const uint64_t ahi = p.a_mn ? make_mnmajor_sw128_desc(a_hi) : make_kmajor_sw128_desc(a_hi);
                    const uint64_t alo = p.a_mn ? make_mnmajor_sw128_desc(a_lo) : make_kmajor_sw128_desc(a_lo);
                    const uint64_t bhi = p.b_mn ? make_mnmajor_sw128_desc(b_hi) : make_kmajor_sw128_desc(b_hi);
                    const uint64_t blo = p.b_mn ? make_mnmajor_sw128_desc(b_lo) : make_kmajor_sw128_desc(b_lo);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t aa = a_adv * (uint64_t)k, ba = b_adv * (uint64_t)k;
                        mma(d_tmem, alo + aa, bhi + ba, (uint32_t)((kb != c0) | (k != 0)));
                        mma(d_tmem, ahi + aa, blo + ba, 1u);
                        mma(d_tmem, ahi + aa, bhi + ba, 1u);
                    }
                    commit(empty_bar(s));              // smem stage is free (in both CTAs) once these MMAs retire
                    if constexpr (WIDE) commit(loempty_bar(l));
                    if (++s == kS) { s = 0; ph ^= 1u; }
                    if (++l == kL) { l = 0; lph ^= 1u; }
                }
                commit(tfull_bar(as));                 // chunk accumulator complete
                if (++as == 2) { as = 0; aph ^= 1u; }
            }
            stamp(p, 3);                               // last MMA issued
        }
    } else if (kConv > 0 && (warp < kEpi0 || warp >= kEpi0 + kEpiWarps)) {
        // converter warps (in the WIDE layout: warps 2-3 of warpgroup 0 and the whole of warpgroup 3)
        if constexpr (cfg::kWide) {
            if (warp >= kEpi0 + kEpiWarps) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");   // warpgroup 3
        }
        run_converter();
    } else {
        // ===================== epilogue warps 2..9 =====================
        // warp w may touch the TMEM lanes of quadrant (w & 3); the two warps of a quadrant split the
        // BN columns in halves.  Each thread owns ONE output row and HALF = BN/2 consecutive columns,
        // and keeps their running sum over the K chunks of a segment in registers.
        if constexpr (cfg::kWide) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        constexpr int HALF = BN / 2;
        const int quad = warp & 3;
        const int half = (warp - kEpi0) >> 2;
        const int trow = quad * 32 + lane;          // row inside the tile
        int* s_flag = reinterpret_cast<int*>(smem_gen + cfg::kTileBytes + kSlotOff + 8);
        int as = 0;
        uint32_t aph = 0;
        int i = 0;
        const bool stamper = threadIdx.x == 32 * kEpi0;
        while (i < my_units) {
            int tile, ch0;
            unit_at(p, i, dp_units, sk_begin, nb, bid, tile, ch0);
            // the segment: this CTA's consecutive units of `tile`
            const int seg_len = i < dp_units ? p.cpt : min(p.cpt - ch0, my_units - i);
            const int m0 = (tile / p.tiles_n) * TM + (int)cta_rank * BM, n0 = (tile % p.tiles_n) * BN;
            float acc[HALF];
#pragma unroll
            for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
            for (int e = i + seg_len; i < e; ++i) {
                mbar_wait(tfull_bar(as), aph);
                tcgen05_fence_after();
#pragma unroll
                for (int c = 0; c < HALF; c += 32) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + half * HALF + c);
                    tmem_ld_32x32(taddr, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c + j] += __uint_as_float(r[j]);  // IEEE RN adds across chunks
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_leader(tempty_bar(as), pair_leader);
                    else mbar_arrive(tempty_bar(as));
                }
                if (++as == 2) { as = 0; aph ^= 1u; }
            }
            if (stamper && i == my_units) stamp(p, 4);   // last accumulator chunk drained
            if (seg_len == p.cpt) {
                // the segment is the whole tile: add the bias and store
                const int row = m0 + trow;
                const int col0 = n0 + half * HALF;
                if (row < p.m) {
                    const float bv = (p.bias != nullptr) ? p.bias[row] : 0.f;
                    if (p.ep_mode == GEMM_EP_NONE) {
#pragma unroll
                        for (int c = 0; c < HALF; c += 4) store_quad(p, row, col0 + c, acc[c] + bv, acc[c + 1] + bv, acc[c + 2] + bv, acc[c + 3] + bv);
                    } else {
                        const EpArgs ea = ep_args(p);
#pragma unroll
                        for (int c = 0; c < HALF; c += 16)
                            store16_ep(ea, row, col0 + c, make_float4(acc[c] + bv, acc[c + 1] + bv, acc[c + 2] + bv, acc[c + 3] + bv),
                                       make_float4(acc[c + 4] + bv, acc[c + 5] + bv, acc[c + 6] + bv, acc[c + 7] + bv),
                                       make_float4(acc[c + 8] + bv, acc[c + 9] + bv, acc[c + 10] + bv, acc[c + 11] + bv),
                                       make_float4(acc[c + 12] + bv, acc[c + 13] + bv, acc[c + 14] + bv, acc[c + 15] + bv));
                    }
                }
            } else if (p.kcluster > 1) {
                // ---- split-K inside a cluster: exchange the partial tile through distributed shared memory ----
                // The K ranges of this tile are the CTAs (pairs) of this cluster.  Thread t of every CTA holds the same
                // row and the same HALF columns; the columns are cut into kcluster parts and part j is finished by the
                // CTA with K index j.  Every CTA stores its other parts straight into the finisher's shared memory
                // (the operand stages are free once every CTA's MMAs have retired: first cluster barrier), the second
                // cluster barrier publishes them, and the finisher adds the contributions in K order (deterministic,
                // the same order as the workspace scheme) and stores its part of C from registers.
                auto exchange = [&](auto nseg_c) {
                    constexpr int NSEG = decltype(nseg_c)::value;
                    constexpr int NC4 = HALF / 4;
                    constexpr int P4 = NC4 / NSEG > 0 ? NC4 / NSEG : 1;     // float4 per part and thread
                    const int t = threadIdx.x - 32 * kEpi0;
                    cluster_sync_all();                                    // every CTA of the cluster has drained its last accumulator chunk
#pragma unroll
                    for (int j = 0; j < NSEG; ++j) {
                        if (j == (int)kidx) continue;
                        const uint32_t target = PAIR ? (uint32_t)(2 * j) + cta_rank : (uint32_t)j;
#pragma unroll
                        for (int c = 0; c < P4; ++c) {
                            const int a0 = (j * P4 + c) * 4;
                            st_dsmem_f4(smem_base + 16u * (uint32_t)(((int)kidx * P4 + c) * 256 + t), target,
                                        make_float4(acc[a0], acc[a0 + 1], acc[a0 + 2], acc[a0 + 3]));
                        }
                    }
                    cluster_sync_all();                                    // all contributions have landed
                    const int row = m0 + trow;
                    const float bv = (p.bias != nullptr && row < p.m) ? p.bias[row] : 0.f;
                    const float4* recv = reinterpret_cast<const float4*>(smem_gen) + t;
#pragma unroll
                    for (int j = 0; j < NSEG; ++j) {
                        if (j != (int)kidx) continue;                       // j == kidx: the part this CTA finishes (compile-time register indices)
#pragma unroll
                        for (int c = 0; c < P4; ++c) {
                            const int a0 = (j * P4 + c) * 4;
                            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int src = 0; src < NSEG; ++src) {
                                const float4 v = (src == j) ? make_float4(acc[a0], acc[a0 + 1], acc[a0 + 2], acc[a0 + 3]) : recv[(src * P4 + c) * 256];
                                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                            }
                            if (row < p.m) store_quad(p, row, n0 + half * HALF + a0, sum.x + bv, sum.y + bv, sum.z + bv, sum.w + bv);
                        }
                    }
                };
                if (p.kcluster == 2) exchange(std::integral_constant<int, 2>());
                else if (p.kcluster == 4) exchange(std::integral_constant<int, 4>());
                else exchange(std::integral_constant<int, 8>());
            } else {
                // partial segment: park it in this CTA's slot (0 = the segment that starts the CTA's range)
                const int sk_tile = tile - p.dp_tiles;
                const int which = (sk_tile == sk_begin / p.cpt) ? 0 : 1;
                if (p.perm_s > 1) {
                    // aligned split-K: plane-major slot, float4 index = (c / 4) * 256 + epilogue thread — every store
                    // instruction of a warp writes 512 contiguous bytes (the row-major slot below writes 16 bytes to
                    // each of 32 rows: parking 64 KB took 3.4 us that way)
                    float4* slot = reinterpret_cast<float4*>(p.slots + (((size_t)lb * 2 + which) * (PAIR ? 2 : 1) + cta_rank) * (BM * BN)) + (threadIdx.x - 32 * kEpi0);
#pragma unroll
                    for (int c = 0; c < HALF; c += 4) slot[(c / 4) * 256] = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
                } else {
                    float* slot = p.slots + (((size_t)lb * 2 + which) * (PAIR ? 2 : 1) + cta_rank) * (BM * BN) + (size_t)trow * BN + half * HALF;
#pragma unroll
                    for (int c = 0; c < HALF; c += 4) *reinterpret_cast<float4*>(slot + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
                }
                __threadfence();
                epi_bar();
                const int b_first = unit_owner(p, sk_tile * p.cpt), b_last = unit_owner(p, (sk_tile + 1) * p.cpt - 1);
                unsigned int* ticket = p.tickets + (PAIR ? 2 * tile + (int)cta_rank : tile);   // the two halves of a pair's tile are finished independently
                if (p.perm_s > 1) {
                    // Aligned split-K: every worker of the tile holds exactly ONE segment and all of them
                    // finish together (one CTA per SM, grid <= SMs: co-resident), so instead of leaving
                    // the whole tile to the last arriver (a serial tail of nseg + 1 tile-sized passes on
                    // one SM: measured 6.8 us for 4 x 128 x 256, 15.9 us when the slots miss L2) the CTAs wait for
                    // each other and each adds 1/nseg of the rows: segments in K order (deterministic).
                    const unsigned nseg = (unsigned)(b_last - b_first + 1);
                    if (threadIdx.x == 32 * kEpi0) {
                        stamp(p, 6);   // segment parked
                        // the counter runs 0 .. 2 nseg - 1 and wraps to 0 by itself (atomicInc): nseg arrivals, then nseg
                        // departures; neither needs its return value, so nothing waits for the round trip
                        atomicInc(ticket, 2u * nseg - 1u);
                        unsigned int seen;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ticket) : "memory");
                        } while (seen < nseg);
                        stamp(p, 7);   // all segments of the tile parked
                    }
                    epi_bar();
                    __threadfence();
                    const int t = threadIdx.x - 32 * kEpi0;
                    constexpr int NC4 = HALF / 4;                                  // float4 planes of a slot
                    const int mine = lb - b_first;                                 // this worker's position among the tile's segments
                    const int c4_0 = (int)((unsigned)mine * NC4 / nseg), c4_1 = (int)((unsigned)(mine + 1) * NC4 / nseg);
                    const float4* seg0 = reinterpret_cast<const float4*>(p.slots + (((size_t)b_first * 2) * (PAIR ? 2 : 1) + cta_rank) * (BM * BN)) + t;
                    const size_t seg_stride = (size_t)2 * (PAIR ? 2 : 1) * (BM * BN) / 4;   // float4 between the slots of consecutive workers
                    // thread t reduces what thread t of every segment parked: its own row, columns half * HALF + 4 c4 ...
                    const int row = m0 + trow;
                    const float bv = (p.bias != nullptr && row < p.m) ? p.bias[row] : 0.f;
                    for (int c4 = c4_0; c4 < c4_1; c4 += 8) {
                        float4 sum[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) sum[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (unsigned j0 = 0; j0 < nseg; j0 += 2) {     // 2 segments x 8 planes in flight
                            float4 v[2][8];
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    v[jj][q] = (j0 + jj < nseg && c4 + q < c4_1) ? __ldcg(seg0 + (j0 + jj) * seg_stride + (size_t)(c4 + q) * 256)
                                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    if (j0 + jj < nseg) { sum[q].x += v[jj][q].x; sum[q].y += v[jj][q].y; sum[q].z += v[jj][q].z; sum[q].w += v[jj][q].w; }
                                }
                        }
                        if (row < p.m) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (c4 + q >= c4_1) continue;
                                store_quad(p, row, n0 + half * HALF + (c4 + q) * 4, sum[q].x + bv, sum[q].y + bv, sum[q].z + bv, sum[q].w + bv);
                            }
                        }
                    }
                    epi_bar();   // every thread has read the segments
                    if (threadIdx.x == 32 * kEpi0) atomicInc(ticket, 2u * nseg - 1u);   // the last departure wraps the counter to 0
                    continue;
                }
                if (threadIdx.x == 32 * kEpi0) *s_flag = (atomicAdd(ticket, 1u) == (unsigned)(b_last - b_first));
                epi_bar();
                const bool last = (*s_flag != 0);
                epi_bar();   // everyone has read the flag before the next segment overwrites it
                if (last) {
                    // last arriver: C tile = sum of the tile's segments in K order (+ bias).  The
                    // register accumulators are dead here; the tile is walked as float4 index
                    // f = it * 256 + t (fully coalesced), 8 float4 per thread per pass with the loads
                    // of 2 segments in flight together.
                    __threadfence();
                    const int t = threadIdx.x - 32 * kEpi0;
                    constexpr int F4_PER_ROW = BN / 4;
                    constexpr int PASSES = BM * BN / 4 / 256 / 8;
                    for (int pass = 0; pass < PASSES; ++pass) {
                        float4 sum[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) sum[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int b0 = b_first; b0 <= b_last; b0 += 2) {
                            float4 v[2][8];
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int b = b0 + j;
                                if (b <= b_last) {
                                    const int w = (sk_tile == unit_begin(p, b) / p.cpt) ? 0 : 1;
                                    const float4* src = reinterpret_cast<const float4*>(p.slots + (((size_t)b * 2 + w) * (PAIR ? 2 : 1) + cta_rank) * (BM * BN));
#pragma unroll
                                    for (int q = 0; q < 8; ++q) v[j][q] = __ldcg(src + (pass * 8 + q) * 256 + t);
                                } else {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) v[j][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 2; ++j)
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    sum[q].x += v[j][q].x; sum[q].y += v[j][q].y; sum[q].z += v[j][q].z; sum[q].w += v[j][q].w;
                                }
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int f = (pass * 8 + q) * 256 + t;
                            const int r = m0 + f / F4_PER_ROW, cc = n0 + (f % F4_PER_ROW) * 4;
                            if (r < p.m) {
                                const float bv = (p.bias != nullptr) ? p.bias[r] : 0.f;
                                store_quad(p, r, cc, sum[q].x + bv, sum[q].y + bv, sum[q].z + bv, sum[q].w + bv);
                            }
                        }
                    }
                    if (threadIdx.x == 32 * kEpi0) *ticket = 0;
                }
            }
        }
    }

    if (threadIdx.x == 32 * kEpi0) stamp(p, 5);          // epilogue done (stores, fix-up)
    if (p.kcluster > 1 && !(warp >= kEpi0 && warp < kEpi0 + kEpiWarps)) {
        cluster_sync_all();   // the two cluster barriers of the split-K exchange (the epilogue warps pass them inside the exchange)
        cluster_sync_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();   // neither CTA may leave (or free TMEM) while the pair's MMAs / barrier arrivals are in flight
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::kTmemCols) : "memory");
    }
}

// ---- split pre-pass ---------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// what the tensor core makes of a raw fp32 word under kind::tf32: the low 13 mantissa bits are ignored
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// lo plane for an operand whose hi part is the source itself
__device__ __forceinline__ float lo_of(float x) {
    const float res = x - tf32_trunc(x);  // exact in fp32; 0 for +-inf after the isfinite test below
    return isfinite(res) ? tf32_rna(res) : 0.f;
}
// general route: both planes are written, hi rounded to nearest
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    float h = tf32_rna(x);
    if (!isfinite(h)) h = tf32_trunc(x);  // rounding overflowed: truncate instead
    hi = h;
    const float res = x - h;  // exact in fp32
    lo = isfinite(res) ? tf32_rna(res) : 0.f;
}

struct SplitJob {
    const float* src;
    float* hi;            // null: the tensor core reads `src` in place (mode 0)
    float* lo;
    int rows;             // extent of the operand's M / N index
    int mode;             // 0: lo only, same layout as src (count = rows * k elements, multiple of 4)
                          // 1: src is [rows x k] (K contiguous) -> padded planes [rows x kp]
                          // 2: src is [k x rows]                 -> padded planes [rows x kp], transposed
    int blocks;           // CTAs assigned to this operand
};

// One launch splits BOTH operands of a GEMM (blockIdx.x < ja.blocks -> operand A, else B).
__global__ void __launch_bounds__(256) split_pair_kernel(const SplitJob ja, const SplitJob jb, int k, int kp) {
    __shared__ float tile[32][33];
    const bool second = blockIdx.x >= (unsigned)ja.blocks;
    const SplitJob& j = second ? jb : ja;
    const int bid = second ? blockIdx.x - ja.blocks : blockIdx.x;
    if (j.mode == 0) {
        const size_t nvec = (size_t)j.rows * k / 4;
        const size_t base = (size_t)bid * 1024 + threadIdx.x;
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t i = base + u * 256;
            if (i < nvec) v[u] = reinterpret_cast<const float4*>(j.src)[i];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t i = base + u * 256;
            if (i < nvec) reinterpret_cast<float4*>(j.lo)[i] = make_float4(lo_of(v[u].x), lo_of(v[u].y), lo_of(v[u].z), lo_of(v[u].w));
        }
    } else if (j.mode == 1) {
        const int vec_in = ((k % 4) == 0) && ((reinterpret_cast<uintptr_t>(j.src) & 15) == 0);
        const int kq = kp / 4;
        const size_t total = (size_t)j.rows * kq;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t i = (size_t)bid * 1024 + u * 256 + threadIdx.x;
            if (i >= total) break;
            const int r = (int)(i / kq), c = (int)(i % kq) * 4;
            float x[4] = {0.f, 0.f, 0.f, 0.f};
            if (vec_in && c + 4 <= k) {
                const float4 v = *reinterpret_cast<const float4*>(j.src + (size_t)r * k + c);
                x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (c + e < k) x[e] = j.src[(size_t)r * k + c + e];
            }
            float4 h, l;
            split_tf32(x[0], h.x, l.x); split_tf32(x[1], h.y, l.y); split_tf32(x[2], h.z, l.z); split_tf32(x[3], h.w, l.w);
            *reinterpret_cast<float4*>(j.hi + (size_t)r * kp + c) = h;
            *reinterpret_cast<float4*>(j.lo + (size_t)r * kp + c) = l;
        }
    } else {
        const int tiles_r = (j.rows + 31) / 32;
        const int r0 = (bid % tiles_r) * 32, k0 = (bid / tiles_r) * 32;
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
        for (int e = ty; e < 32; e += 8) {
            const int kk = k0 + e, rr = r0 + tx;
            tile[e][tx] = (kk < k && rr < j.rows) ? j.src[(size_t)kk * j.rows + rr] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int e = ty; e < 32; e += 8) {
            const int rr = r0 + e, kk = k0 + tx;
            if (rr < j.rows && kk < kp) {
                float h, l;
                split_tf32(tile[tx][e], h, l);
                j.hi[(size_t)rr * kp + kk] = h;
                j.lo[(size_t)rr * kp + kk] = l;
            }
        }
    }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
    if (g_encode) return DSC_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return fail(DSC_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return DSC_OK;
}

// 2-D map over a row-major fp32 array [outer x inner] (row pitch `pitch` elements); box = 32 floats
// (128 B, one swizzle row) x box_outer rows, SWIZZLE_128B, out-of-range elements read as zero.
//   K-major operand : inner = K, outer = M/N rows, box_outer = tile rows
//   MN-major operand: inner = M/N, outer = K,      box_outer = BK (one box per 32-wide MN block),
//                     SWIZZLE_128B_ATOM_32B (the only layout the tensor core takes for 32-bit MN-major)
static int make_map(CUtensorMap* map, const float* base, int inner, int outer, size_t pitch, int box_outer, bool mn_major = false) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DSC_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d array, box rows %d", (int)r, outer, inner, box_outer);
    return DSC_OK;
}

// `workers` = CTAs, or CTA pairs (clusters of 2: the two SMs of a TPC) when PAIR
template <int BN, bool PAIR, bool WIDE>
static int set_smem_attr() {
    using cfg = Cfg<BN, PAIR, WIDE>;
    static bool attr_set = false;
    if (!attr_set) {
        DSC_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, PAIR, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::kSmemBytes));
        attr_set = true;
    }
    return DSC_OK;
}

template <int BN, bool PAIR, bool WIDE = false>
static int launch(Ctx* ctx, const CUtensorMap* maps, const Params& p, int workers) {
    using cfg = Cfg<BN, PAIR, WIDE>;
    DSC_TRY((set_smem_attr<BN, PAIR, WIDE>()));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(PAIR ? 2 * workers : workers);
    lc.blockDim = dim3(cfg::kThreads);
    lc.dynamicSmemBytes = cfg::kSmemBytes;
    lc.stream = ctx->stream;
    const int cluster = (PAIR ? 2 : 1) * (p.kcluster > 1 ? p.kcluster : 1);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = cluster > 1 ? 1 : 0;
    DSC_CUDA_CHECK(cudaLaunchKernelEx(&lc, gemm_tf32x3_kernel<BN, PAIR, WIDE>, maps[0], maps[1], maps[2], maps[3], p));
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

// How many clusters of `cluster` CTAs of this kernel can be resident at once (0 on error); cached per (kernel, size).
template <int BN, bool PAIR, bool WIDE = false>
static int max_active_clusters(int cluster) {
    using cfg = Cfg<BN, PAIR, WIDE>;
    static int cache[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
    if (cluster < 1 || cluster > 8) return 0;
    if (cache[cluster] >= 0) return cache[cluster];
    if (set_smem_attr<BN, PAIR, WIDE>() != DSC_OK) return 0;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(cluster * 64);
    lc.blockDim = dim3(cfg::kThreads);
    lc.dynamicSmemBytes = cfg::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tf32x3_kernel<BN, PAIR, WIDE>, &lc) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    cache[cluster] = n;
    return n;
}

// How one operand reaches the tensor core.
struct OperandPlan {
    bool direct;       // hi = the source buffer in place; only a lo plane is written
    bool mn_major;     // tiles are MN-major (direct operands whose K index is strided)
    int mode;          // SplitJob mode
    size_t ws_floats;  // workspace floats (lo plane, or hi + lo padded planes)
};

// src_k_contig: the stored array is [mn x k]; otherwise [k x mn]
static OperandPlan plan_operand(const float* src, bool src_k_contig, int mn, int k, int kp) {
    OperandPlan o;
    const int contig = src_k_contig ? k : mn;
    static const bool force_planes = [] { const char* e = getenv("DSC_GEMM_PLANES"); return e && *e == '1'; }();
    o.direct = !force_planes && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (contig % 4) == 0;
    o.mn_major = o.direct && !src_k_contig;
    o.mode = o.direct ? 0 : (src_k_contig ? 1 : 2);
    o.ws_floats = o.direct ? (size_t)mn * k : 2 * (size_t)mn * kp;
    return o;
}

static int split_blocks(const OperandPlan& o, int rows, int k, int kp) {
    if (o.mode == 0) return (int)(((size_t)rows * k / 4 + 1023) / 1024);
    if (o.mode == 1) return (int)(((size_t)rows * (kp / 4) + 1023) / 1024);
    return ((rows + 31) / 32) * (kp / 32);
}

// ---- the planner: a pure function of the shape, the SM count and the pinned choices -------------------
// (exported as dsc_gemm_plan so that the schedule arithmetic is unit-tested on a machine without a GPU)
struct PlanIn {
    int sm_count, m, n, k, num_kb, cpt;
    bool a_direct, b_direct;      // operand is addressable by TMA in place (plan_operand)
    bool a_ready, b_ready;        // ... and its operand-split companion already exists (written by the producing kernel, or by an earlier product)
    int fuse;                     // 0: never split in the kernel; 1: single-CTA kernels may; 2: also the WIDE pair kernel
    int force_bn, force_pair, force_split;
    bool allow_wide;
};
struct PlanOut {
    int bn, grid, dp_waves, dp_extra, sk_units_total, perm_s;
    bool pair;
};
static bool make_plan(const PlanIn& in, PlanOut& out) {
    int bn = 256, grid = 1, dp_waves = 0, dp_extra = 0, sk_units_total = 0, perm_s = 0;
    bool pair = false;
    const int cpt = in.cpt;
    const double kb_per_chunk = (double)in.num_kb / cpt;
    const int m = in.m, n = in.n, k = in.k;
    const int fuse_env = in.fuse, force_bn = in.force_bn, force_pair = in.force_pair, force_split = in.force_split;
    const bool allow_wide = in.allow_wide;
    auto has_converters = [allow_wide](int c, bool pr) { return pr ? (allow_wide && c == 256) : c < 256; };
    {
        // Per worker (CTA, or CTA pair) and 32-wide k-block, in cycles:
        //   tensor pipe : 4 k-slices x 3 MMAs x (BN/256 x 128)                               = 6 BN
        //   smem        : TMA writes + MMA operand reads (3 MMAs re-read both operands) [+ converter read/write],
        //                 over the rows one CTA stages (128 + BN, pair: 128 + BN/2), 128 B/cycle  = 5 (6) x rows
        //   L2 -> SM    : the TMA bytes of every active CTA through ~6 KB/cycle
        double best = -1.0;
        for (int cand = 0; cand < 6; ++cand) {
            const int c = cand % 3 == 0 ? 256 : (cand % 3 == 1 ? 128 : 64);
            const bool pr = cand >= 3;
            if (force_bn && c != force_bn) continue;
            if (force_pair >= 0 && (int)pr != force_pair) continue;
            const int tm = pr ? 2 * BM : BM;
            const int tiles_m_c = (m + tm - 1) / tm;
            const long tiles = (long)tiles_m_c * ((n + c - 1) / c);
            if (tiles * cpt > 0x3fffffff) continue;
            const int workers = pr ? in.sm_count / 2 : in.sm_count;
            const int ctas_per_worker = pr ? 2 : 1;
            const int brows = pr ? c / 2 : c;
            const bool can_fuse = fuse_env && has_converters(c, pr);
            bool fa = can_fuse && in.a_direct && !in.a_ready, fb = can_fuse && in.b_direct && !in.b_ready;   // a ready companion is simply loaded
            if (pr) fa = fb = (fa && fb);   // the pair kernel splits both operands in the kernel or neither
            const double bytes_kb = (double)BM * (fa ? 128.0 : 256.0) + (double)brows * (fb ? 128.0 : 256.0);
            const double smem_kb = (double)BM * (fa ? 6.0 : 5.0) + (double)brows * (fb ? 6.0 : 5.0);
            // pre-pass for the operands that are not split in the kernel: 8 bytes per element at ~5 TB/s + launch
            const double pre_elems = ((fa || in.a_ready) ? 0.0 : (double)m * k) + ((fb || in.b_ready) ? 0.0 : (double)n * k);
            const double pre = pre_elems == 0.0 ? 0.0 : 2000.0 * (3.0 + pre_elems * 8.0 / 5.0e6);
            auto per_kb = [&](double active_workers) {
                double t = 6.0 * c;
                if (smem_kb > t) t = smem_kb;
                const double l2 = bytes_kb * active_workers * ctas_per_worker / 6000.0;
                if (l2 > t) t = l2;
                return t;
            };
            const double fix_scale = (double)c / 256.0;   // cost of parking / re-reading partial tiles scales with the tile width
            auto take = [&](double cost, int g, int waves, int extra, long sk, int ps) {
                if (best < 0 || cost < best) {
                    best = cost; bn = c; pair = pr; grid = g; dp_waves = waves; dp_extra = extra; sk_units_total = (int)sk; perm_s = ps;
                }
            };
            if (tiles >= workers) {
                const int g = workers;
                const int waves = (int)(tiles / g);
                const int r = (int)(tiles % g);
                const double chunk_t = kb_per_chunk * per_kb(g);
                // (a) plain last wave: r workers run one more whole tile
                take(((double)waves + (r ? 1.0 : 0.0)) * cpt * chunk_t + 4000.0 + pre, g, waves, r, 0, 0);
                // (b) stream-K over the last wave.  Nearly every worker parks two segments and re-reads
                // two or three: extra L2 traffic, all at the same time (measured: ~15 us for 148 CTAs
                // with 128 x 256 tiles) — pays only when K is deep.
                if (r) {
                    const long sk = (long)r * cpt;
                    const long sk_per = (sk + g - 1) / g;
                    take(((double)waves * cpt + (double)sk_per) * chunk_t + 4000.0 + 30000.0 * fix_scale + pre, g, waves, 0, sk, 0);
                }
            } else {
                const int smax = cpt < 16 ? cpt : 16;
                for (int sp = 1; sp <= smax; ++sp) {
                    if (tiles * sp > workers) break;
                    if (force_split > 0 && sp != force_split && !(sp == smax && force_split > smax)) continue;
                    const int g = (int)tiles * sp;
                    const long per_w = (cpt + sp - 1) / sp;
                    // park + ticket + the last arriver's pass over sp segments (a serial tail)
                    double cost = (double)per_w * kb_per_chunk * per_kb(g) + 4000.0 + (sp > 1 ? (6000.0 + 2500.0 * sp) * fix_scale : 0.0) + pre;
                    if (cpt % sp != 0) cost *= 1.15;  // ranges cross tile boundaries: lockstep is lost
                    take(cost, g, 0, 0, tiles * cpt, (sp > 1 && cpt % sp == 0) ? sp : 0);
                }
            }
        }
        if (best < 0) return false;
    }
    out.bn = bn; out.pair = pair; out.grid = grid; out.dp_waves = dp_waves; out.dp_extra = dp_extra;
    out.sk_units_total = sk_units_total; out.perm_s = perm_s;
    return true;
}

}  // namespace tf32x3

int gemm_tf32x3_tcgen05(Ctx* ctx, int ta, int tb, int m, int n, int k, MatRef Am, MatRef Bm, const float* bias, float* C, const GemmEpilogue* ep) {
    using namespace tf32x3;
    const float* A = reinterpret_cast<const float*>(Am.ptr);
    const float* B = reinterpret_cast<const float*>(Bm.ptr);
    // Eligibility: enough rows/cols to fill 128-row tiles and enough depth to amortise the pre-pass.
    // Skinny / tiny products stay on the CUDA-core kernel (they are bandwidth- or latency-bound).
    if (m < 128 || n < 32 || k < 32) return 1;
    DSC_TRY(load_encode());
    const int kp = (k + BK - 1) / BK * BK;

    // ---- plan: tile width, grid and schedule (all decided before anything is launched) ----
    // Candidates: tile width 256 / 128 / 64, one CTA per tile (128 rows) or a CTA pair (256 rows); the
    // cycle model is stated next to the candidate loop.
    // Schedules (one kernel, see Params):
    //   tiles >= SMs : whole-tile waves (neighbouring CTAs walk K in lockstep over neighbouring tiles,
    //                  so operand panels are shared in L2) + stream-K over the tiles of the last,
    //                  partial wave, whose operands fit in L2 whatever the order
    //   tiles <  SMs : aligned split-K — every tile is cut into `s` equal K ranges (grid = tiles * s);
    //                  CTAs of the same range stay in lockstep.  Unaligned stream-K ranges would keep
    //                  all SMs busy but every CTA would stream its own operand slice from HBM
    //                  (measured: 1024 x 1024 x 8192 ran 50 % slower that way).
    const int num_kb = kp / BK;
    const int cpt = (num_kb + kChunkKb - 1) / kChunkKb;
    const int tiles_m = (m + BM - 1) / BM;
    // op(A)(m x k): stored m x k (K contiguous) unless ta; op(B)(k x n): stored n x k (K contiguous) only if tb
    const OperandPlan pa = plan_operand(A, !ta, m, k, kp), pb = plan_operand(B, tb != 0, n, k, kp);
    // In-kernel split (tiles narrower than 256): a direct operand needs no pre-pass and no lo plane;
    // only its fp32 words cross L2 -> SM (half the bytes), and the converter warps derive the lo tile.
    static const int fuse_env0 = [] { const char* e = getenv("DSC_GEMM_FUSE"); return e ? atoi(e) : 1; }();
    static const int force_bn0 = [] { const char* e = getenv("DSC_GEMM_BN"); return e ? atoi(e) : 0; }();
    const int fuse_env = ctx->gemm_force_fuse >= 0 ? ctx->gemm_force_fuse : fuse_env0;
    const int force_bn = ctx->gemm_force_bn > 0 ? ctx->gemm_force_bn : force_bn0;
    static const int force_pair0 = [] { const char* e = getenv("DSC_GEMM_PAIR"); return e ? atoi(e) : -1; }();   // -1: planner decides
    const int force_pair = ctx->gemm_force_pair >= 0 ? ctx->gemm_force_pair : force_pair0;
    const int force_split = ctx->gemm_force_split;
    // Kernel variants with converter warps (Cfg::kConvWarps > 0).  The WIDE pair kernel is correct but
    // measured slower than pre-pass + plain pair kernel on every shape tried (8192 x 1024 x 1024:
    // 91.5 us against 80.6 us including the pre-pass): the in-kernel split repeats the conversion for
    // every tile that re-reads an operand and its 64 KB of converter traffic per k-block competes with
    // the MMA operand fetch for shared-memory bandwidth.  It is only planned when asked for
    // (dsc_gemm_tune fuse = 2), which is how the parity tests and tools/gemm_sweep.py reach it.
    const bool allow_wide = ctx->gemm_force_fuse == 2;
    auto has_converters = [allow_wide](int c, bool pr) { return pr ? (allow_wide && c == 256) : c < 256; };
    PlanIn pin;
    pin.sm_count = (ctx->gemm_sm_budget > 0 && ctx->gemm_sm_budget < ctx->sm_count) ? ctx->gemm_sm_budget : ctx->sm_count;
    if (!ctx->on_lane && ctx->main_gemm_half_once) {
        // the second compute lane has just been given a product planned for half of the SMs (its partner of the same tape node):
        // this one takes the other half if the lane is still running it
        ctx->main_gemm_half_once = false;
        if (lane_busy(ctx) && ctx->gemm_sm_budget == 0) pin.sm_count = ctx->sm_count / 2;
    }
    pin.m = m; pin.n = n; pin.k = k; pin.num_kb = num_kb; pin.cpt = cpt;
    pin.a_direct = pa.direct; pin.b_direct = pb.direct;
    // Operand-split companions (Block::lo): present when the kernel that produced the matrix wrote one next to it (bias +
    // activation, adjoint chains, GEMM epilogues) or an earlier product split it and nobody has written the block since.
    auto companion = [&](const MatRef& M, size_t count, bool direct) -> float* {
        Block* blk = M.block;
        if (!direct || !blk || !blk->lo || blk->lo_epoch != ctx->cache_epoch || M.off < blk->lo_begin || M.off + count > blk->lo_end) return nullptr;
        if (lane_gate_read(ctx, blk->lo) != DSC_OK) return nullptr;   // a companion the second compute lane wrote
        return reinterpret_cast<float*>(blk->lo->ptr) + M.off;
    };
    float* a_comp = companion(Am, (size_t)m * k, pa.direct);
    float* b_comp = companion(Bm, (size_t)n * k, pb.direct);
    pin.a_ready = a_comp != nullptr; pin.b_ready = b_comp != nullptr;
    pin.fuse = fuse_env; pin.force_bn = force_bn; pin.force_pair = force_pair; pin.force_split = force_split;
    pin.allow_wide = allow_wide;
    PlanOut pout;
    if (!make_plan(pin, pout)) return 1;
    const int bn = pout.bn, grid = pout.grid, dp_waves = pout.dp_waves, dp_extra = pout.dp_extra, sk_units_total = pout.sk_units_total, perm_s = pout.perm_s;
    const bool pair = pout.pair;
    Params p;
    p.m = m; p.n = n; p.num_kb = num_kb;
    p.tiles_m = pair ? (m + 2 * BM - 1) / (2 * BM) : tiles_m;
    p.tiles_n = (n + bn - 1) / bn;
    p.cpt = cpt;
    const long tiles = (long)p.tiles_m * p.tiles_n;
    p.dp_waves = dp_waves;
    p.dp_tiles = dp_waves * grid;
    p.dp_extra = dp_extra;
    p.perm_s = perm_s;
    p.perm_tiles = perm_s ? (int)tiles : 0;
    p.base = sk_units_total / grid;
    p.rem = sk_units_total % grid;
    static const bool debug_plan = [] { const char* e = getenv("DSC_GEMM_DEBUG"); return e && *e == '1'; }();
    bool a_fused = fuse_env && has_converters(bn, pair) && pa.direct && !a_comp, b_fused = fuse_env && has_converters(bn, pair) && pb.direct && !b_comp;
    if (pair) a_fused = b_fused = (a_fused && b_fused);
    if (debug_plan)
        fprintf(stderr, "[dsc gemm] m=%d n=%d k=%d ta=%d tb=%d: BN=%d pair=%d grid=%d cpt=%d dp_waves=%d dp_extra=%d sk_units=%d perm_s=%d fused=%d%d ready=%d%d ep=%d\n", m, n, k, ta, tb, bn,
                (int)pair, grid, cpt, dp_waves, dp_extra, sk_units_total, perm_s, (int)a_fused, (int)b_fused, (int)(a_comp != nullptr), (int)(b_comp != nullptr), ep ? ep->mode : 0);
    // Aligned split-K whose K ranges fit one cluster (<= 8 CTAs) and whose clusters are all resident at once:
    // the partial tiles go through distributed shared memory (no workspace round trip, no tickets).
    p.kcluster = 0;
    {
        static const int kc_env = [] { const char* e = getenv("DSC_GEMM_KCLUSTER"); return e ? atoi(e) : 1; }();
        const int sp = perm_s;
        const bool wide = pair && bn == 256 && a_fused && b_fused;
        if (kc_env && !wide && (sp == 2 || sp == 4) && (pair ? 2 : 1) * sp <= 4 && (bn / 8) % sp == 0) {   // clusters of 2 and 4 CTAs (8: not exercised on hardware yet)
            const int cl = (pair ? 2 : 1) * sp;
            int fit = 0;
            if (pair) fit = bn == 256 ? max_active_clusters<256, true>(cl) : (bn == 128 ? max_active_clusters<128, true>(cl) : max_active_clusters<64, true>(cl));
            else fit = bn == 256 ? max_active_clusters<256, false>(cl) : (bn == 128 ? max_active_clusters<128, false>(cl) : max_active_clusters<64, false>(cl));
            if (fit >= (int)tiles) {
                p.kcluster = sp;
                p.perm_tiles = 0;   // worker = tile * sp + K index: the K ranges of a tile are consecutive CTAs (one cluster)
            }
            if (debug_plan) fprintf(stderr, "[dsc gemm]   split-K through a cluster of %d: %d clusters fit, %ld tiles -> %s\n", cl, fit, tiles, p.kcluster ? "yes" : "no");
        }
    }
    DSC_TRY(ensure_gemm_tickets(ctx, (size_t)tiles * (pair ? 2 : 1)));
    p.tickets = ctx->gemm_tickets;

    // Where each operand's lo plane lives: a ready companion; a NEW companion attached to the operand's block (direct operands:
    // the split below fills it and later products of the same matrix find it); or the context workspace (operands TMA cannot
    // address in place: hi + lo padded planes).
    auto new_companion = [&](const MatRef& M, size_t count) -> float* {
        Block* blk = M.block;
        if (!blk) return nullptr;
        if (!blk->lo) {
            Block* lo = nullptr;
            if (pool_alloc(ctx, blk->bytes, &lo) != DSC_OK) return nullptr;
            blk->lo = lo;
        } else if (lane_gate_write(ctx, blk->lo) != DSC_OK) {
            return nullptr;
        }
        blk->lo_epoch = ctx->cache_epoch;
        blk->lo_begin = M.off;
        blk->lo_end = M.off + count;
        return reinterpret_cast<float*>(blk->lo->ptr) + M.off;
    };
    const bool a_split = !a_fused && !a_comp, b_split = !b_fused && !b_comp;   // this call runs the split for the operand
    if (a_split && pa.direct) a_comp = new_companion(Am, (size_t)m * k);
    if (b_split && pb.direct) b_comp = new_companion(Bm, (size_t)n * k);
    const size_t a_floats = (a_fused || a_comp) ? 0 : (pa.ws_floats + 255) & ~size_t(255), b_floats = (b_fused || b_comp) ? 0 : (pb.ws_floats + 255) & ~size_t(255);
    const size_t planes_bytes = (a_floats + b_floats) * sizeof(float);
    const size_t slot_bytes = (size_t)grid * (pair ? 2 : 1) * 2 * BM * bn * sizeof(float);
    DSC_TRY(ensure_gemm_ws(ctx, planes_bytes + slot_bytes + 2048));
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ctx->gemm_ws) + 1023) & ~uintptr_t(1023));
    float* ws = reinterpret_cast<float*>(base);
    // direct: [lo]; planes: [hi | lo]
    float* a_lo = a_comp ? a_comp : (pa.direct ? ws : ws + (size_t)m * kp);
    float* b_lo = b_comp ? b_comp : (pb.direct ? ws + a_floats : ws + a_floats + (size_t)n * kp);
    const float* a_hi = pa.direct ? A : ws;
    const float* b_hi = pb.direct ? B : ws + a_floats;
    p.slots = reinterpret_cast<float*>(base + planes_bytes);

    // ---- (1) split pre-pass: only for the operands that have no lo plane yet and that the kernel does not split itself ----
    if (a_split || b_split) {
        SplitJob ja{A, pa.direct ? nullptr : ws, a_lo, m, pa.mode, a_split ? split_blocks(pa, m, k, kp) : 0};
        SplitJob jb{B, pb.direct ? nullptr : ws + a_floats, b_lo, n, pb.mode, b_split ? split_blocks(pb, n, k, kp) : 0};
        split_pair_kernel<<<ja.blocks + jb.blocks, 256, 0, ctx->stream>>>(ja, jb, k, kp);
        DSC_LAUNCH_CHECK(ctx);
    }

    // ---- (2) tensor-core kernel ----
    p.C = C;
    p.bias = bias;
    p.ep_mode = GEMM_EP_NONE; p.ep_op = -1; p.ep_vec = nullptr; p.ep_aux = nullptr; p.ep_out2 = nullptr; p.ep_lo = nullptr;
    bool ep_in_kernel = false;
    if (ep && ep->mode != GEMM_EP_NONE) {
        // The kernel can apply bias rows (+ Tanh / Sigmoid / ReL) and the three adjoint chains to the finished tile.  Whether it
        // SHOULD is a measured decision (dsc_gemm_epilogue): the tile is drained by 8 warps per SM, one thread per row, and
        // tanhf / coshf per element make those warps — not the tensor pipe — the bottleneck (8192 x 1024 x 1024 + bias + tanh:
        // 115-125 us fused against 75 + 14.5 us as two kernels, the second one HBM-bound with 64 warps per SM).  Default
        // (mode 1): only the cheap epilogues run in the kernel — bias rows without activation, ReL and its adjoint chain;
        // mode 2: all of them (parity tests); mode 0: none.  Re-measured in round 2 on the one-wave products of a batch
        // shard (1024 x 1024 x 1024, at most one tile per CTA): bias + tanh in the kernel 26.4 us against 19.8 + 4.2 us,
        // the tanh adjoint chain 38.7 us against 25.5 + 4.4 us — the transcendental epilogues stay separate kernels.
        const int mode = ctx->gemm_epilogue;
        const bool cheap = (ep->mode == GEMM_EP_BIAS_ACT && (ep->op < 0 || ep->op == DSC_OP_REL)) || (ep->mode == GEMM_EP_CHAIN && ep->op == DSC_FUSED_RELU_BWD);
        const bool able = ep->mode == GEMM_EP_CHAIN || ep->op < 0 || ep->op == DSC_OP_TANH || ep->op == DSC_OP_SIGMOID || ep->op == DSC_OP_REL;
        if (able && (mode == 2 || (mode == 1 && cheap))) {
            ep_in_kernel = true;
            p.ep_mode = ep->mode; p.ep_op = ep->op;
            p.ep_vec = reinterpret_cast<const float*>(ep->vec);
            p.ep_aux = reinterpret_cast<const float*>(ep->aux);
            p.ep_out2 = reinterpret_cast<float*>(ep->out2);
            p.ep_lo = ep->lo;
        }
    }
    p.vec4 = ((n % 4) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    p.a_mn = pa.mn_major ? 1 : 0;
    p.b_mn = pb.mn_major ? 1 : 0;
    p.a_fused = a_fused ? 1 : 0;
    p.b_fused = b_fused ? 1 : 0;
    static const bool want_stamps = [] { const char* e = getenv("DSC_GEMM_STAMPS"); return e && *e == '1'; }();
    if (want_stamps && !ctx->gemm_stamps && !ctx->capturing) {
        DSC_CUDA_CHECK(cudaMalloc(&ctx->gemm_stamps, 8 * sizeof(unsigned long long)));
        DSC_CUDA_CHECK(cudaMemset(ctx->gemm_stamps, 0, 8 * sizeof(unsigned long long)));
    }
    p.stamps = ctx->gemm_stamps;
    if (a_fused) a_lo = const_cast<float*>(A);   // the lo maps are never dereferenced for a fused operand: point them at valid memory
    if (b_fused) b_lo = const_cast<float*>(B);
    CUtensorMap maps[4];
    if (pa.mn_major) {          // stored k x m
        DSC_TRY(make_map(&maps[0], a_hi, m, k, (size_t)m, BK, true));
        DSC_TRY(make_map(&maps[1], a_lo, m, k, (size_t)m, BK, true));
    } else {                    // [m x k] in place, or padded planes [m x kp]
        const int kk = pa.direct ? k : kp;
        DSC_TRY(make_map(&maps[0], a_hi, kk, m, (size_t)kk, BM));
        DSC_TRY(make_map(&maps[1], a_lo, kk, m, (size_t)kk, BM));
    }
    if (pb.mn_major) {          // stored k x n
        DSC_TRY(make_map(&maps[2], b_hi, n, k, (size_t)n, BK, true));
        DSC_TRY(make_map(&maps[3], b_lo, n, k, (size_t)n, BK, true));
    } else {
        const int kk = pb.direct ? k : kp;
        DSC_TRY(make_map(&maps[2], b_hi, kk, n, (size_t)kk, pair ? bn / 2 : bn));
        DSC_TRY(make_map(&maps[3], b_lo, kk, n, (size_t)kk, pair ? bn / 2 : bn));
    }
    int s;
    if (pair) {
        if (bn == 256 && a_fused && b_fused) s = launch<256, true, true>(ctx, maps, p, grid);
        else if (bn == 256) s = launch<256, true>(ctx, maps, p, grid);
        else if (bn == 128) s = launch<128, true>(ctx, maps, p, grid);
        else s = launch<64, true>(ctx, maps, p, grid);
    } else {
        if (bn == 256) s = launch<256, false>(ctx, maps, p, grid);
        else if (bn == 128) s = launch<128, false>(ctx, maps, p, grid);
        else s = launch<64, false>(ctx, maps, p, grid);
    }
    if (s != DSC_OK) return s;
    ctx->stats.gemm_tcgen05++;
    return (ep && ep->mode != GEMM_EP_NONE && !ep_in_kernel) ? 2 : DSC_OK;   // 2: product stored, the caller applies the epilogue as a second pass
}

}  // namespace dsc

extern "C" int dsc_gemm_plan(int sm_count, int m, int n, int k, int a_direct, int b_direct, int bn, int pair, int split, int fuse, int32_t out[20]) {
    using namespace dsc;
    using namespace dsc::tf32x3;
    if (!out || sm_count < 2 || m < 0 || n < 0 || k < 0) return fail(DSC_E_INVALID, "dsc_gemm_plan: bad arguments");
    for (int i = 0; i < 20; ++i) out[i] = 0;
    if (m < 128 || n < 32 || k < 32) return DSC_OK;   // not eligible: out[0] = 0
    const int kp = (k + BK - 1) / BK * BK;
    PlanIn pin;
    pin.sm_count = sm_count; pin.m = m; pin.n = n; pin.k = k;
    pin.num_kb = kp / BK;
    pin.cpt = (pin.num_kb + kChunkKb - 1) / kChunkKb;
    pin.a_direct = a_direct != 0; pin.b_direct = b_direct != 0;
    pin.a_ready = a_direct == 2; pin.b_ready = b_direct == 2;   // 2: direct and its operand-split companion exists already
    pin.fuse = fuse < 0 ? 1 : fuse;
    pin.force_bn = bn > 0 ? bn : 0;
    pin.force_pair = pair < 0 ? -1 : (pair != 0);
    pin.force_split = split;
    pin.allow_wide = fuse == 2;
    PlanOut po;
    if (!make_plan(pin, po)) return DSC_OK;
    const int tiles_m = po.pair ? (m + 2 * BM - 1) / (2 * BM) : (m + BM - 1) / BM;
    const int tiles_n = (n + po.bn - 1) / po.bn;
    const long tiles = (long)tiles_m * tiles_n;
    auto has_converters = [&](int c, bool pr) { return pr ? (pin.allow_wide && c == 256) : c < 256; };
    bool a_fused = pin.fuse && has_converters(po.bn, po.pair) && pin.a_direct && !pin.a_ready, b_fused = pin.fuse && has_converters(po.bn, po.pair) && pin.b_direct && !pin.b_ready;
    if (po.pair) a_fused = b_fused = (a_fused && b_fused);
    const bool wide = po.pair && po.bn == 256 && a_fused && b_fused;
    const int sp = po.perm_s;
    const int kc = (!wide && (sp == 2 || sp == 4) && (po.pair ? 2 : 1) * sp <= 4 && (po.bn / 8) % sp == 0) ? sp : 0;
    out[0] = 1; out[1] = po.bn; out[2] = po.pair ? 1 : 0; out[3] = po.grid; out[4] = pin.cpt; out[5] = pin.num_kb;
    out[6] = tiles_m; out[7] = tiles_n; out[8] = po.dp_waves; out[9] = po.dp_extra; out[10] = po.sk_units_total; out[11] = po.perm_s;
    out[12] = po.sk_units_total / po.grid; out[13] = po.sk_units_total % po.grid; out[14] = po.dp_waves * po.grid;
    out[15] = po.perm_s ? (int)tiles : 0; out[16] = kc; out[17] = a_fused ? 1 : 0; out[18] = b_fused ? 1 : 0;
    return DSC_OK;
}

extern "C" int dsc_gemm_debug_stamps(dsc_ctx_t c, uint64_t out[8]) {
    using namespace dsc;
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (!ctx->gemm_stamps) return fail(DSC_E_INVALID, "no GEMM time stamps: set DSC_GEMM_STAMPS=1 before the first product");
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    DSC_CUDA_CHECK(cudaMemcpy(out, ctx->gemm_stamps, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return DSC_OK;
}

extern "C" int dsc_gemm_epilogue(dsc_ctx_t c, int mode) {
    using namespace dsc;
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    if (mode < 0 || mode > 2) return fail(DSC_E_INVALID, "dsc_gemm_epilogue: mode must be 0, 1 or 2");
    as_ctx(c)->gemm_epilogue = mode;
    return DSC_OK;
}

extern "C" int dsc_gemm_tune(dsc_ctx_t c, int bn, int pair, int split, int fuse) {
    using namespace dsc;
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    if (!(bn == -1 || bn == 256 || bn == 128 || bn == 64)) return fail(DSC_E_INVALID, "dsc_gemm_tune: bn must be -1, 256, 128 or 64");
    Ctx* ctx = as_ctx(c);
    ctx->gemm_force_bn = bn;
    ctx->gemm_force_pair = pair < 0 ? -1 : (pair != 0);
    ctx->gemm_force_split = split;
    ctx->gemm_force_fuse = fuse < 0 ? -1 : (fuse > 2 ? 2 : fuse);
    return DSC_OK;
}
