// Implicit-GEMM contraction kernel for sm_100a: one persistent, warp-specialised kernel that
// serves every Linear / Conv2d 3x3 / Conv2d 3x3 stride 2 / Conv3d (3,1,1) on the SVD UNet and
// temporal-VAE path (reference call sites listed in include/wiw.h).
//
//   warp 0      TMA producer: per k-iteration one 5-D box load of A (a 128-pixel patch x 64 channels,
//               shifted by the filter tap; out-of-range = zero padding) and one 2-D box of W
//   warp 1      MMA issuer: tcgen05.mma kind::f16, M=128 (cta_group::1) or M=256 over a CTA pair (cta_group::2), N=bn,
//               K=16 x4 per stage, fp32 accumulators in TMEM, double-buffered across tiles
//   warp 2      TMEM allocator + TMA loader of the epilogue's residual / aux tiles (one slot ring per epilogue group)
//   warp 3      idle (completes warpgroup 0, which gives its registers away: setmaxnreg 40; the epilogue warps take 232)
//   warps 4-11  epilogue, two groups of 4 warps on alternating 32-column chunks: tcgen05.ld -> (bias,
//               per-row-group table, SiLU, GEGLU, residual blend) -> fp16 -> swizzled smem -> TMA store
//               (clips partial tiles).  Every warp stages and stores its OWN 32-row slab (2 KB, its own
//               bulk-store groups), so the epilogue warps never meet at a CTA barrier.  LEAN = 1 (no activation / aux):
//               two chunks in flight per warp, packed f32x2 arithmetic, bias from a per-group shared-memory copy.
//
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue),
// residual full/empty ring (loader <-> epilogue), bulk-group ring for the stores.
#include "wiw_common.cuh"
#include "../../include/wiw.h"

#include <cstdlib>

namespace wiw {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                    // 64 fp16 = one 128-byte swizzle row
constexpr int kEpiCols = 32;                   // epilogue chunk: 32 fp16 = 64-byte rows (SWIZZLE_64B)
constexpr int kEpiChunkBytes = kBlockM * kEpiCols * 2;   // 8 KB
constexpr int kMaxOutBufs = 8;                 // staging chunks: 2 or 4 per epilogue group (GemmArgs::out_bufs)
constexpr int kMaxResSlots = 8;                // residual / aux chunk slots: GemmArgs::res_slots (4 or 8)
constexpr int kMaxStages = 8;
constexpr int kThreads = 12 * 32;              // warpgroup 0: TMA, MMA, epilogue loader, (idle); warpgroups 1-2: 2 x 4 epilogue warps
constexpr int kBiasStageBytes = 2 * 2 * 512;    // lean epilogue: per epilogue group, two tiles' worth (accumulator-stage parity) of bias values
constexpr int kRegsProducer = 40;              // setmaxnreg budgets: 128 * 40 + 256 * 232 = 64512 = 384 * 168 (what the launch allocates)
constexpr int kRegsEpilogue = 232;
constexpr int kSmemBudget = 227 * 1024;

struct GemmArgs {
    int total_tiles, n_tiles, tiles_x, tiles_y;   // total_tiles counts PAIR tiles when cta_group == 2
    int bx, by, bz, dim_x, dim_y;                 // M-tile = bz images x by rows x bx pixels (bx*by*bz == 128)
    int bn, stages, num_taps, chunks_per_tap;
    int out_cols_per_tile, geglu, has_res, has_aux;
    int tab_div, tab_mod, n_total, tmem_cols;
    int tab_stride, act, out_total;
    int dbg;                                      // development aid (timing only, results garbage): bit 0 = skip TMA loads, bit 1 = skip MMA issue,
                                                  // bit 3 = GEGLU epilogue multiplies value * gate without the GELU, bit 4 = skip staging + TMA stores
    int res_slots;                                // residual (aux) chunks in flight: 8 lets a whole tile's residual prefetch behind the MMAs
    int out_bufs;                                 // staging buffers per epilogue group: 4 when no residual / aux slots compete for smem, else 2
    int warp_store;                               // 1: every epilogue warp stores its own 32-row slab (no CTA barriers); 0: one store per 128-row chunk
    int wx0[4], wy0[4], wz0[4];                   // (pixel, row, image) offset of epilogue warp q's 32 rows inside the M-tile
    const float* bias;
    const __half* tab;
    float s_acc, s_res, s_aux;
    int tap_cbase[9], tap_ox[9], tap_p[9], tap_oy[9];
};

// development aid (WIW_GEMM_DBG bit 2): cycle counters of CTA 0's producer / MMA / first epilogue warp, read by wiw_debug_gemm_prof
__device__ long long g_gemm_prof[24];     // [8..15]: epilogue phases of CTA 0 warp 3 (see eph_* below), [16] chunks

struct TileCoord {
    int x0, y0, img, n_tile;
};
// tile = index of a (pair of) M-tile(s) x N-tile; with cta_group::2 CTA `rank` of the pair owns M-tile 2*mp + rank
// (an M-tile past the end decodes to an out-of-range image: its loads zero-fill and its stores clip)
template <int CG>
__device__ __forceinline__ TileCoord decode_tile(const GemmArgs& g, int tile, int rank) {
    TileCoord t;
    const int mp = tile / g.n_tiles;
    t.n_tile = tile - mp * g.n_tiles;
    const int m_tile = mp * CG + rank;
    int tx = m_tile % g.tiles_x;
    int r = m_tile / g.tiles_x;
    int ty = r % g.tiles_y;
    t.img = (r / g.tiles_y) * g.bz;
    t.x0 = tx * g.bx;
    t.y0 = ty * g.by;
    return t;
}

// ----------------------------------------------------------------------------- lean-epilogue pieces (one 32-column chunk of one row per thread)
// The epilogue warps are ISSUE-bound (two warps per scheduler, ~350 instructions per pair of chunks): all fp32 arithmetic is done
// on register pairs with the packed f32x2 forms (FADD2 / FMUL2 / FFMA2: same pipe time as the scalar forms, half the issue slots --
// scripts/ffma2_bench.cu).  Results are bit-identical to the scalar code: same operations, same rounding, two lanes per instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 half2_to_f32x2(const __half2 h) { const float2 f = __half22float2(h); return pk2(f.x, f.y); }
constexpr int kEpiPairs = kEpiCols / 2;

// bias: 32 floats in shared memory (the group's copy, staged once per tile); tab: this row's 32 fp16 table values, already in registers
__device__ __forceinline__ void lean_bias_tab(f32x2 (&v)[kEpiPairs], const uint32_t (&r)[kEpiCols], const float* bias_smem, bool has_tab,
                                              const uint4 (&tu)[kEpiCols / 8]) {
#pragma unroll
    for (int c = 0; c < kEpiPairs; ++c) v[c] = pk2u(r[2 * c], r[2 * c + 1]);
    if (bias_smem != nullptr) {
#pragma unroll
        for (int c = 0; c < kEpiPairs; c += 2) {
            const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(bias_smem + 2 * c);
            v[c] = add2(v[c], b.x); v[c + 1] = add2(v[c + 1], b.y);
        }
    }
    if (has_tab) {
#pragma unroll
        for (int c = 0; c < kEpiCols / 8; ++c) {
            const __half2* h = reinterpret_cast<const __half2*>(&tu[c]);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[4 * c + e] = add2(v[4 * c + e], half2_to_f32x2(h[e]));
        }
    }
}
// v *= gelu(gate + bias) with mul_gelu_tanhfit_f's polynomial, evaluated in y = x / 2 so that the halving costs nothing: the staged
// gate bias is ALREADY halved (power-of-two scalings commute with rounding, so every intermediate equals the scalar code's):
//   y = 0.5 g + 0.5 b;  tanh argument x p(x^2) = y q(y^2) with q(s) = 2 c0 + 8 c1 s + 32 c2 s^2;  v * 0.5 x = v y
__device__ __forceinline__ void lean_geglu(f32x2 (&v)[kEpiPairs], const uint32_t (&rg)[kEpiCols], const float* half_bias_smem) {
    const f32x2 half = pk2(0.5f, 0.5f);
    const f32x2 q2 = pk2(32.0f * -3.58549415e-4f, 32.0f * -3.58549415e-4f), q1 = pk2(8.0f * 3.70488970e-2f, 8.0f * 3.70488970e-2f),
                q0 = pk2(2.0f * 7.97460674e-1f, 2.0f * 7.97460674e-1f);
#pragma unroll
    for (int c = 0; c < kEpiPairs; c += 2) {
        ulonglong2 hb = make_ulonglong2(0ull, 0ull);
        if (half_bias_smem != nullptr) hb = *reinterpret_cast<const ulonglong2*>(half_bias_smem + 2 * c);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const f32x2 y = fma2(pk2u(rg[2 * (c + e)], rg[2 * (c + e) + 1]), half, e ? hb.y : hb.x);
            const f32x2 y2 = mul2(y, y);
            f32x2 p = fma2(y2, q2, q1);
            p = fma2(p, y2, q0);
            float a0, a1, t0, t1;
            upk2(mul2(y, p), a0, a1);
            asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
            const f32x2 w = mul2(v[c + e], y);
            v[c + e] = fma2(w, pk2(t0, t1), w);
        }
    }
}
// the group's k-th residual chunk sits in its own two-slot ring (lean kernels run with 4 slots): slot grp * 2 + (k & 1), phase (k >> 1) & 1
__device__ __forceinline__ uint32_t lean_residual_load(uint4 (&u)[4], const uint8_t* slots_row, uint32_t grp, uint32_t k, uint32_t bar_rfull, uint32_t sw) {
    const uint32_t rs = grp * 2u + (k & 1u);
    mbar_wait(bar_rfull + 8 * rs, (k >> 1) & 1u);
    const uint8_t* p = slots_row + rs * kEpiChunkBytes;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) u[c4] = *reinterpret_cast<const uint4*>(p + ((uint32_t(c4) ^ sw) << 4));
    return rs;
}
__device__ __forceinline__ void lean_residual_add(f32x2 (&v)[kEpiPairs], const uint4 (&u)[4], float s_res) {
    const f32x2 s2 = pk2(s_res, s_res);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
        const __half2* h = reinterpret_cast<const __half2*>(&u[c4]);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[c4 * 4 + e] = fma2(s2, half2_to_f32x2(h[e]), v[c4 * 4 + e]);
    }
}
__device__ __forceinline__ void lean_stage(uint8_t* p, const f32x2 (&v)[kEpiPairs], uint32_t sw) {
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
        uint4 u;
        __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float a, b;
            upk2(v[c4 * 4 + e], a, b);
            h[e] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<uint4*>(p + ((uint32_t(c4) ^ sw) << 4)) = u;
    }
}

// LEAN = 1 fixes at compile time what most layers of the model have anyway (no activation, no aux operand, per-warp stores,
// two staging buffers, four residual slots, no debug switches): the chunk loop of an epilogue warp is a serial chain of short latencies
// (two warps per scheduler, nothing to hide them behind), and every runtime switch in it costs a constant-bank load + compare +
// branch (profiles/r02_gemm_epilogue_phases_v2.txt: ~280 cycles to add a bias, ~170 to skip the release).
template <int CG, int GEGLU, int LEAN>      // GEGLU: 0 = plain epilogue, 2 = GEGLU epilogue (exact-erf GELU through the one-MUFU tanh fit, mul_gelu_tanhfit_f)
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out /* box = one warp's 32 rows x 32 channels */,
                 const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_aux,
                 const __grid_constant__ GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x (A 16 KB | B bn*128)] [out bufs] [res slots] [aux slots] [barriers]
    // The dynamic window is declared 1024-byte aligned (SWIZZLE_128B atoms) and the host budgets NO slack for a run-time round-up:
    // the 1 KB it used to cost is what keeps the 6th pipeline stage next to the bias copies.  A base that is not aligned traps.
    const uint32_t smem_base = smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0u) asm volatile("trap;");
    const uint32_t a_bytes = kBlockM * kBlockK * 2;
    const uint32_t b_bytes = (g.bn / CG) * kBlockK * 2;          // with cta_group::2 each CTA stages half of the N-tile
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t out_base = smem_base + g.stages * stage_bytes;
    const uint32_t res_base = out_base + 2 * g.out_bufs * kEpiChunkBytes;
    const uint32_t aux_base = res_base + (g.has_res ? g.res_slots * kEpiChunkBytes : 0);
    const uint32_t bar_base = aux_base + (g.has_aux ? g.res_slots * kEpiChunkBytes : 0);
    // barrier slots (8 B each)
    const uint32_t bar_full = bar_base;                         // [kMaxStages]
    const uint32_t bar_empty = bar_full + 8 * kMaxStages;       // [kMaxStages]
    const uint32_t bar_tfull = bar_empty + 8 * kMaxStages;      // [2]
    const uint32_t bar_tempty = bar_tfull + 16;                 // [2]
    const uint32_t bar_rfull = bar_tempty + 16;                 // [kMaxResSlots]
    const uint32_t bar_rempty = bar_rfull + 8 * kMaxResSlots;   // [kMaxResSlots]
    const uint32_t tmem_slot = bar_rempty + 8 * kMaxResSlots;   // uint32
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    // lean epilogue: [group][accumulator stage][128 floats] bias values of the chunks the group handles in a tile (kBiasStageBytes)
    float* const bias_stage = reinterpret_cast<float*>(smem_raw + (((tmem_slot + 16 + 15) & ~15u) - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int k_iters = g.num_taps * g.chunks_per_tap;
    const bool has_epi_loads = g.has_res || g.has_aux;
    const int cta_rank = (CG == 2) ? (int)cluster_ctarank() : 0;
    const int first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        tma_prefetch_desc(&tm_out);
        if (g.has_res) tma_prefetch_desc(&tm_res);
        if (g.has_aux) tma_prefetch_desc(&tm_aux);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < g.stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 8 * CG);       // one arrival per epilogue warp of every CTA in the pair
        }
        for (int i = 0; i < kMaxResSlots; ++i) {
            mbar_init(bar_rfull + 8 * i, 1);
            mbar_init(bar_rempty + 8 * i, 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CG == 2) { tmem_alloc_2sm(tmem_slot, g.tmem_cols); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, g.tmem_cols); tmem_relinquish(); }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // The three producer warps need few registers; the epilogue warps' chunk loop is a serial latency chain that only more
    // independent work per warp (two chunks in flight, held in registers) shortens.  Rebalance the register file by warpgroup.
    // (setmaxnreg sits INSIDE each role's branch: ptxas budgets code reachable from both settings with the smaller one.)
    const uint32_t acc_stride = (g.bn <= 32) ? 32 : (g.bn <= 64 ? 64 : (g.bn <= 128 ? 128 : 256));

    if (warp == 0) {
        // ================================ TMA producer ================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
        // The whole warp runs the loop (warp-uniform control flow) and ONE elected lane issues: a TMA / tcgen05
        // instruction reached under a divergent `lane == 0` branch is wrapped by ptxas in an ELECT + BRA.U.ANY
        // serialisation loop with its operands moved through R2UR, which costs ~100 cycles per instruction
        // (profiles/r01_mma_cost.md); from a converged warp the UTMALDG / UTCHMMA / UTCBAR are issued back to back.
        int s = 0;
        uint32_t ph = 0;
        const bool prof = (g.dbg & 4) && blockIdx.x == 0;
        long long p_wait = 0;
        const long long p_t0 = prof ? clock64() : 0;
        for (int tile = first_tile; tile < g.total_tiles; tile += tile_step) {
            const TileCoord t = decode_tile<CG>(g, tile, cta_rank);
            const int n0 = t.n_tile * g.bn + cta_rank * (g.bn / CG);
            int kcol = 0;
            for (int tap = 0; tap < g.num_taps; ++tap) {
                const int cb = g.tap_cbase[tap], cx = t.x0 + g.tap_ox[tap], cp = g.tap_p[tap], cy = t.y0 + g.tap_oy[tap];
                for (int ch = 0; ch < g.chunks_per_tap; ++ch) {
                    const long long w0 = prof ? clock64() : 0;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    if (prof) p_wait += clock64() - w0;
                    const uint32_t sa = smem_base + s * stage_bytes;
                    if (elect_one()) {
                        if (g.dbg & 1) {
                            if (cta_rank == 0) mbar_arrive(bar_full + 8 * s);
                        } else if (CG == 2) {
                            // both CTAs' bytes are credited to the leader's barrier; only the leader arms it
                            if (cta_rank == 0) mbar_expect_tx(bar_full + 8 * s, 2 * stage_bytes);
                            tma_load_5d_2sm(sa, &tm_a, bar_full + 8 * s, cb + ch * kBlockK, cx, cp, cy, t.img);
                            tma_load_2d_2sm(sa + a_bytes, &tm_b, bar_full + 8 * s, kcol, n0);
                        } else {
                            mbar_expect_tx(bar_full + 8 * s, stage_bytes);
                            tma_load_5d(sa, &tm_a, bar_full + 8 * s, cb + ch * kBlockK, cx, cp, cy, t.img);
                            tma_load_2d(sa + a_bytes, &tm_b, bar_full + 8 * s, kcol, n0);
                        }
                    }
                    __syncwarp();
                    kcol += kBlockK;
                    if (++s == g.stages) { s = 0; ph ^= 1; }
                }
            }
        }
        if (prof && lane == 0) { g_gemm_prof[4] = clock64() - p_t0; g_gemm_prof[5] = p_wait; }
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
        // ================================ MMA issuer ================================
        if (cta_rank == 0) {                              // the leader CTA issues the MMAs of the pair
            const uint32_t idesc = make_idesc_f16(g.bn, 0, 0, 128 * CG);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            const bool prof = (g.dbg & 4) && blockIdx.x == 0;
            long long m_wfull = 0, m_wacc = 0, m_steps = 0;
            const long long m_t0 = prof ? clock64() : 0;
            for (int tile = first_tile; tile < g.total_tiles; tile += tile_step) {
                const long long w0 = prof ? clock64() : 0;
                mbar_wait(bar_tempty + 8 * as, aph ^ 1);
                if (prof) m_wacc += clock64() - w0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * acc_stride;
                for (int it = 0; it < k_iters; ++it) {
                    const long long w1 = prof ? clock64() : 0;
                    mbar_wait(bar_full + 8 * s, ph);
                    if (prof) { m_wfull += clock64() - w1; ++m_steps; }
                    tc_fence_after();
                    const uint32_t sa = smem_base + s * stage_bytes;
                    const uint64_t da = make_desc_sw128(sa);
                    const uint64_t db = make_desc_sw128(sa + a_bytes);
                    if (elect_one()) {                    // converged warp, one issuing lane (see the producer's note)
                        if (!(g.dbg & 2)) {
#pragma unroll
                            for (int kk = 0; kk < kBlockK / 16; ++kk) {
                                // +32 bytes per K=16 step inside the 128-byte swizzle row (units of 16 B)
                                if (CG == 2) umma_f16_ss_2sm(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                                else umma_f16_ss(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                            }
                        }
                        // frees the smem stage (in both CTAs of a pair) when the MMAs retire
                        if (CG == 2) umma_commit_2sm(bar_empty + 8 * s, 3); else umma_commit(bar_empty + 8 * s);
                        if (it == k_iters - 1) {          // accumulator ready
                            if (CG == 2) umma_commit_2sm(bar_tfull + 8 * as, 3); else umma_commit(bar_tfull + 8 * as);
                        }
                    }
                    __syncwarp();
                    if (++s == g.stages) { s = 0; ph ^= 1; }
                }
                if (++as == 2) { as = 0; aph ^= 1; }
            }
            if (prof && lane == 0) { g_gemm_prof[0] = clock64() - m_t0; g_gemm_prof[1] = m_wfull; g_gemm_prof[2] = m_wacc; g_gemm_prof[3] = m_steps; }
        }
    } else if (warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
        // ================================ epilogue loader ================================
        if (has_epi_loads) {
            const int full_chunks = g.out_cols_per_tile / kEpiCols;
            const uint32_t bytes = kEpiChunkBytes * ((g.has_res ? 1 : 0) + (g.has_aux ? 1 : 0));
            // Each epilogue group owns HALF of the slots as its own ring (group of chunk j = j & 1; the group's k-th chunk sits in its
            // slot k % per, phase (k / per) & 1).  With one ring shared by both groups, consecutive uses of a slot could belong to
            // different groups (odd chunk counts per tile): a group that ran ahead then waited on the slot's barrier one phase early,
            // the parity test passed on the phase before, and it read the slot before its data had landed (seen as rare wrong residual
            // values once the epilogue got faster; scripts/stress_gemm_res.py).  One consumer group per slot makes its uses ordered.
            const uint32_t per = uint32_t(g.res_slots) >> 1, per_shift = (per == 4u) ? 2u : 1u;
            uint32_t cnt[2] = {0u, 0u};
            for (int tile = first_tile; tile < g.total_tiles; tile += tile_step) {
                const TileCoord t = decode_tile<CG>(g, tile, cta_rank);
                const int c0 = t.n_tile * g.out_cols_per_tile;
                const int chunks = min(full_chunks, (g.out_total - c0) / kEpiCols);     // the last N-tile may be ragged
                for (int j = 0; j < chunks; ++j) {
                    const uint32_t gsel = uint32_t(j) & 1u;
                    const uint32_t k = gsel ? cnt[1]++ : cnt[0]++;
                    const uint32_t rs = gsel * per + (k & (per - 1u));
                    const uint32_t rph = (k >> per_shift) & 1u;
                    mbar_wait(bar_rempty + 8 * rs, rph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_rfull + 8 * rs, bytes);
                        if (g.has_res) tma_load_4d(res_base + rs * kEpiChunkBytes, &tm_res, bar_rfull + 8 * rs, c0 + j * kEpiCols, t.x0, t.y0, t.img);
                        if (g.has_aux) tma_load_4d(aux_base + rs * kEpiChunkBytes, &tm_aux, bar_rfull + 8 * rs, c0 + j * kEpiCols, t.x0, t.y0, t.img);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));      // idle: completes warpgroup 0 for setmaxnreg
    } else {
        // ================================ epilogue (warps 4..11, two groups of 4) ================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
        // Group grp handles the 32-column chunks j with j % 2 == grp; both groups read the same TMEM lanes
        // (a warp may only touch lane quadrant warp % 4).  Two warps per SM sub-partition hide each other's
        // TMEM-load / barrier / store latency.
        const int grp = (warp - 4) >> 2;
        const int q = warp & 3;                        // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;                 // row of the 128-row tile == TMEM lane
        const bool store_warp = (((warp - 4) & 3) == 0);   // its elected lane owns the group's bulk-store groups (elect.sync is deterministic)
        const int full_chunks = g.out_cols_per_tile / kEpiCols;
        const uint32_t sw = (uint32_t(row) >> 1) & 3u;  // SWIZZLE_64B: 16-B chunk index ^= (row>>1)&3
        const uint32_t row_off = uint32_t(row) * 64u;
        const uint32_t bar_id = 1 + grp;
        int as = 0;
        uint32_t aph = 0;
        int ob = 0;                                    // this group's staging buffer (g.out_bufs per group, round robin)
        const int nob = LEAN ? 2 : g.out_bufs;
        const bool e_tab = g.tab != nullptr;
        const int e_act = LEAN ? 0 : g.act;
        const float s_acc = g.s_acc;
        const bool e_sacc = s_acc != 1.0f;
        const int res_slots = LEAN ? 4 : g.res_slots;
        const uint32_t res_per = uint32_t(res_slots) >> 1, res_per_shift = (res_per == 4u) ? 2u : 1u;       // this group's ring: slots grp * per ...
        const bool e_aux = !LEAN && g.has_aux;
        const bool e_res = g.has_res;
        const bool e_loads = e_res || e_aux;
        const int e_dbg = LEAN ? (g.dbg & 4) : g.dbg;      // bit 4 = profile counters, read once per kernel
        const bool e_warp_store = LEAN ? true : (g.warp_store != 0);
        const bool e_bias = g.bias != nullptr;
        const float s_res = g.s_res;
        const int wx0 = g.wx0[q], wy0 = g.wy0[q], wz0 = g.wz0[q];
        uint32_t res_k = 0;                            // residual / aux chunks this GROUP has consumed so far (its ring position, see the loader)
        int chunks = 0;
        const bool eprof = (e_dbg & 4) && blockIdx.x == 0 && warp == 4;
        long long e_wait = 0;
        // Per-phase cycle sums of the chunk loop (bias add | math | release | residual | wait_read | sts+fence | tma issue | tmem ld+wait)
        // cost ~20 registers, so they exist only in a -DWIW_GEMM_PHASES build (WIW_BUILD_VARIANT=phases, scripts/perf_epilogue_phases.py).
#ifdef WIW_GEMM_PHASES
        long long eph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, ech = 0, ets = 0;      // [8], [9]: first use of the first / second accumulator chunk (lean loop)
#define EPH(i) do { if (eprof) { const long long _n = clock64(); eph[i] += _n - ets; ets = _n; } } while (0)
#define EPH_BEGIN() do { if (eprof) { ets = clock64(); ++ech; } } while (0)
#else
#define EPH(i) do { } while (0)
#define EPH_BEGIN() do { } while (0)
#endif
        const long long e_t0 = eprof ? clock64() : 0;
        for (int tile = first_tile; tile < g.total_tiles; tile += tile_step) {
            const TileCoord t = decode_tile<CG>(g, tile, cta_rank);
            const int acc_n0 = t.n_tile * g.bn;                       // first GEMM column of this tile
            const int out_c0 = t.n_tile * g.out_cols_per_tile;        // first output channel
            chunks = min(full_chunks, (g.out_total - out_c0) / kEpiCols);      // ragged last N-tile: padded columns are skipped
            const __half* tab_row = nullptr;
            if (e_tab) {
                const int x = t.x0 + (row % g.bx), y = t.y0 + (row / g.bx) % g.by, z = t.img + row / (g.bx * g.by);
                long long grow = ((long long)z * g.dim_y + y) * g.dim_x + x;
                if (grow < 0) grow = 0;
                int gi = int((grow / g.tab_div) % g.tab_mod);
                tab_row = g.tab + (size_t)gi * g.tab_stride;
            }
            // Bias values of this group's chunks: lane l fetches float4 (l & 7) of slot (l >> 3) -- plain tiles, slot ci = chunk
            // grp + 2 ci (<= 4 per group); GEGLU tiles (<= 2 chunks per group), slots 0-1 = value-half, slots 2-3 = gate-half bias of
            // chunks grp, grp + 2.  The load is issued before the wait for the accumulator (its latency is never on the chunk loop's
            // chain) and stored after it: the four warps of a group write the SAME values to the group's copy for this accumulator
            // stage, and a full barrier phase on that stage means every warp has finished reading the copy of two tiles ago.
            float* const wbias = bias_stage + (grp * 2 + as) * 128;
            float4 bias_v = make_float4(0, 0, 0, 0);
            if (LEAN && e_bias) {
                const int ci = lane >> 3, part = lane & 7;
                const int j = GEGLU ? grp + 2 * (ci & 1) : grp + 2 * ci;
                const int off = (GEGLU && ci >= 2) ? (g.bn >> 1) : 0;
                if (j < chunks) bias_v = __ldg(reinterpret_cast<const float4*>(g.bias + acc_n0 + off + j * kEpiCols) + part);
                if (GEGLU && ci >= 2) { bias_v.x *= 0.5f; bias_v.y *= 0.5f; bias_v.z *= 0.5f; bias_v.w *= 0.5f; }      // lean_geglu works in x / 2
            }
            const long long e0 = (eprof) ? clock64() : 0;
            mbar_wait(bar_tfull + 8 * as, aph);
            if (eprof) e_wait += clock64() - e0;
            tc_fence_after();
            if (LEAN && e_bias) {
                *reinterpret_cast<float4*>(wbias + lane * 4) = bias_v;
                __syncwarp();
            }
            const uint32_t t_acc = tmem_base + as * acc_stride + (uint32_t(q * 32) << 16);
            const int last_j = ((chunks - 1 - grp) >= 0) ? (grp + 2 * ((chunks - 1 - grp) / 2)) : -1;
            if (last_j < 0) {                          // nothing to do for this group in this tile
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (CG == 2) mbar_arrive_cluster(bar_tempty + 8 * as, 0); else mbar_arrive(bar_tempty + 8 * as); }
            }
            if (LEAN) {
                // ---- lean epilogue: TWO chunks of this group (j0, j0 + 2) in flight per iteration.  An epilogue warp's chunk is a
                // serial chain (tcgen05.ld -> wait -> bias -> [GELU] -> residual -> st.shared -> proxy fence -> bulk store) with one
                // other warp on its scheduler to hide it; two chunks held in registers (setmaxnreg gives the room) overlap the
                // TMEM loads, share ONE fence and ONE bulk-group commit, and halve the per-chunk cost of the serial parts.
                const int half_n = g.bn >> 1;
                uint8_t* const stage = smem_raw + (out_base - smem_u32(smem_raw)) + (grp * 2) * kEpiChunkBytes + row_off;
                for (int j0 = grp; j0 < chunks; j0 += 4) {
                    EPH_BEGIN();
                    const int j1 = j0 + 2;
                    const bool two = j1 < chunks;
                    uint32_t ra[kEpiCols], rb[kEpiCols];
                    f32x2 va[kEpiPairs], vb[kEpiPairs];
                    tmem_ld_32x32b_x32(t_acc + j0 * kEpiCols, ra);
                    if (two) tmem_ld_32x32b_x32(t_acc + j1 * kEpiCols, rb);
                    const int ci0 = (j0 - grp) >> 1;               // bias slot of chunk j0 (j1: the next one)
                    const bool has_tab = !GEGLU && e_tab;
                    uint4 ta[kEpiCols / 8], tb[kEpiCols / 8];
                    if (has_tab) {                                 // per-row table values: issued under the TMEM load
#pragma unroll
                        for (int c = 0; c < kEpiCols / 8; ++c) {
                            ta[c] = __ldg(reinterpret_cast<const uint4*>(tab_row + acc_n0 + j0 * kEpiCols) + c);
                            if (two) tb[c] = __ldg(reinterpret_cast<const uint4*>(tab_row + acc_n0 + j1 * kEpiCols) + c);
                        }
                    }
                    tmem_ld_wait();
                    EPH(7);
#ifdef WIW_GEMM_PHASES
                    if (eprof) {       // how long until the loaded registers can actually be read (the wait above does not stall the clock read)
                        if (ra[0] == 0x7fc12345u) ++ech;
                        EPH(8);
                        if (two && rb[31] == 0x7fc12345u) ++ech;
                        EPH(9);
                    }
#endif
                    lean_bias_tab(va, ra, e_bias ? wbias + ci0 * kEpiCols : nullptr, has_tab, ta);
                    if (two) lean_bias_tab(vb, rb, e_bias ? wbias + (ci0 + 1) * kEpiCols : nullptr, has_tab, tb);
                    EPH(0);
                    if (GEGLU) {
                        tmem_ld_32x32b_x32(t_acc + half_n + j0 * kEpiCols, ra);
                        if (two) tmem_ld_32x32b_x32(t_acc + half_n + j1 * kEpiCols, rb);
                        tmem_ld_wait();
                        lean_geglu(va, ra, e_bias ? wbias + 2 * kEpiCols : nullptr);
                        if (two) lean_geglu(vb, rb, e_bias ? wbias + 3 * kEpiCols : nullptr);
                    }
                    EPH(1);
                    if (j0 + 4 >= chunks) {
                        // this group's TMEM reads of the accumulator stage are done: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(bar_tempty + 8 * as, 0); else mbar_arrive(bar_tempty + 8 * as); }
                    }
                    if (e_sacc) {
                        const f32x2 sa = pk2(s_acc, s_acc);
#pragma unroll
                        for (int c = 0; c < kEpiPairs; ++c) { va[c] = mul2(va[c], sa); vb[c] = mul2(vb[c], sa); }
                    }
                    EPH(2);
                    if (e_res) {
                        const uint8_t* gen = smem_raw + (res_base - smem_u32(smem_raw)) + row_off;
                        uint4 ua[4], ub[4];
                        const uint32_t rsa = lean_residual_load(ua, gen, uint32_t(grp), res_k++, bar_rfull, sw);
                        uint32_t rsb = 0;
                        if (two) rsb = lean_residual_load(ub, gen, uint32_t(grp), res_k++, bar_rfull, sw);
                        // The slots were read through the generic proxy and will be overwritten through the async proxy (the loader's next
                        // TMA load): that order needs a proxy fence before the release.  Without it ~1-3 % of launches picked up a few
                        // 16-byte pieces of the NEXT chunk's residual (scripts/stress_gemm_res3.py; the slower general epilogue hid it).
                        fence_proxy_async_smem();
                        mbar_arrive(bar_rempty + 8 * rsa);
                        if (two) mbar_arrive(bar_rempty + 8 * rsb);
                        lean_residual_add(va, ua, s_res);
                        if (two) lean_residual_add(vb, ub, s_res);
                    }
                    EPH(3);
                    // both staging buffers of the group are rewritten: every earlier bulk store of this warp must have read its slab
                    if (elect_one()) tma_store_wait_read<0>();
                    __syncwarp();
                    EPH(4);
                    lean_stage(stage, va, sw);
                    if (two) lean_stage(stage + kEpiChunkBytes, vb, sw);
                    fence_proxy_async_smem();
                    EPH(5);
                    __syncwarp();                                  // the slabs are complete and visible to the async proxy
                    if (elect_one()) {
                        const uint32_t src = out_base + (grp * 2) * kEpiChunkBytes + q * (32 * 64);
                        tma_store_4d(&tm_out, src, out_c0 + j0 * kEpiCols, t.x0 + wx0, t.y0 + wy0, t.img + wz0);
                        if (two) tma_store_4d(&tm_out, src + kEpiChunkBytes, out_c0 + j1 * kEpiCols, t.x0 + wx0, t.y0 + wy0, t.img + wz0);
                        tma_store_commit();
                    }
                    __syncwarp();
                    EPH(6);
                }
            } else
            for (int j = grp; j < chunks; j += 2) {
                EPH_BEGIN();
                float v[kEpiCols];
                uint32_t rg[kEpiCols];                              // gate half of a GEGLU tile
                const int half_n = g.bn >> 1;
                {
                    // Global loads of bias / table rows are issued BEFORE tcgen05.wait::ld so their latency overlaps the
                    // TMEM load's; the GEGLU gate load is issued before the value half is consumed.
                    uint32_t r[kEpiCols];
                    tmem_ld_32x32b_x32(t_acc + j * kEpiCols, r);
                    const int nb = acc_n0 + j * kEpiCols;
                    float4 b4[kEpiCols / 4];
                    uint4 tu[kEpiCols / 8];
                    if (e_bias) {
#pragma unroll
                        for (int c = 0; c < kEpiCols / 4; ++c) b4[c] = __ldg(reinterpret_cast<const float4*>(g.bias + nb) + c);
                    }
                    if (!GEGLU && e_tab) {
#pragma unroll
                        for (int c = 0; c < kEpiCols / 8; ++c) tu[c] = __ldg(reinterpret_cast<const uint4*>(tab_row + nb) + c);
                    }
                    if (LEAN) {
                        // the staging buffer this chunk will overwrite was read by the store issued two chunks ago: waiting for
                        // that read here, under the TMEM load's latency, takes it off the serial chain
                        if (elect_one()) tma_store_wait_read<1>();
                        __syncwarp();
                    }
                    tmem_ld_wait();
                    EPH(7);
                    if (GEGLU) tmem_ld_32x32b_x32(t_acc + half_n + j * kEpiCols, rg);
#pragma unroll
                    for (int c = 0; c < kEpiCols; ++c) v[c] = __uint_as_float(r[c]);
                    if (e_bias) {
#pragma unroll
                        for (int c = 0; c < kEpiCols; c += 4) {
                            v[c] += b4[c / 4].x; v[c + 1] += b4[c / 4].y; v[c + 2] += b4[c / 4].z; v[c + 3] += b4[c / 4].w;
                        }
                    }
                    if (!GEGLU && e_tab) {
#pragma unroll
                        for (int c = 0; c < kEpiCols; c += 8) {
                            const __half2* h = reinterpret_cast<const __half2*>(&tu[c / 8]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(h[e]);
                                v[c + 2 * e] += f.x; v[c + 2 * e + 1] += f.y;
                            }
                        }
                    }
                    if (!GEGLU && e_act == 1) {
#pragma unroll
                        for (int c = 0; c < kEpiCols; ++c) v[c] = silu_f(v[c]);
                    } else if (!GEGLU && e_act == 2) {           // exact-erf GELU (CLIP ViT-H MLP, transformers ACT2FN["gelu"])
#pragma unroll
                        for (int c = 0; c < kEpiCols; ++c) v[c] = mul_gelu_tanhfit_f(1.0f, v[c]);
                    } else if (!GEGLU && e_act == 3) {           // quick GELU x * sigmoid(1.702 x) (CLIPVisionConfig default)
#pragma unroll
                        for (int c = 0; c < kEpiCols; ++c) v[c] = quick_gelu_f(v[c]);
                    }
                }
                EPH(0);
                if (GEGLU) {
                    const int nb = acc_n0 + half_n + j * kEpiCols;
                    float4 b4[kEpiCols / 4];
#pragma unroll
                    for (int c = 0; c < kEpiCols / 4; ++c)
                        b4[c] = e_bias ? __ldg(reinterpret_cast<const float4*>(g.bias + nb) + c) : make_float4(0, 0, 0, 0);
                    tmem_ld_wait();
                    if (e_dbg & 8) {
#pragma unroll
                        for (int c = 0; c < kEpiCols; ++c) v[c] *= __uint_as_float(rg[c]);
                    } else
#pragma unroll
                    for (int c = 0; c < kEpiCols; c += 4) {
                        v[c] = mul_gelu_tanhfit_f(v[c], __uint_as_float(rg[c]) + b4[c / 4].x);
                        v[c + 1] = mul_gelu_tanhfit_f(v[c + 1], __uint_as_float(rg[c + 1]) + b4[c / 4].y);
                        v[c + 2] = mul_gelu_tanhfit_f(v[c + 2], __uint_as_float(rg[c + 2]) + b4[c / 4].z);
                        v[c + 3] = mul_gelu_tanhfit_f(v[c + 3], __uint_as_float(rg[c + 3]) + b4[c / 4].w);
                    }
                }
                EPH(1);
                if (j == last_j) {
                    // this group's TMEM reads of the accumulator stage are done: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (CG == 2) mbar_arrive_cluster(bar_tempty + 8 * as, 0); else mbar_arrive(bar_tempty + 8 * as); }
                }
                if (e_sacc) {
#pragma unroll
                    for (int c = 0; c < kEpiCols; ++c) v[c] *= s_acc;
                }
                EPH(2);
                if (e_loads) {
                    const int rs = int(uint32_t(grp) * res_per + (res_k & (res_per - 1u)));
                    const uint32_t rph = (res_k >> res_per_shift) & 1u;
                    ++res_k;
                    mbar_wait(bar_rfull + 8 * rs, rph);
                    const uint8_t* gen = smem_raw + (res_base - smem_u32(smem_raw));
                    if (e_res) {
                        const uint8_t* p = gen + rs * kEpiChunkBytes + row_off;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const uint4 u = *reinterpret_cast<const uint4*>(p + ((uint32_t(c4) ^ sw) << 4));
                            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(h[e]);
                                v[c4 * 8 + 2 * e] += s_res * f.x;
                                v[c4 * 8 + 2 * e + 1] += s_res * f.y;
                            }
                        }
                    }
                    if (e_aux) {
                        const uint8_t* p = smem_raw + (aux_base - smem_u32(smem_raw)) + rs * kEpiChunkBytes + row_off;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const uint4 u = *reinterpret_cast<const uint4*>(p + ((uint32_t(c4) ^ sw) << 4));
                            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(h[e]);
                                v[c4 * 8 + 2 * e] += g.s_aux * f.x;
                                v[c4 * 8 + 2 * e + 1] += g.s_aux * f.y;
                            }
                        }
                    }
                    fence_proxy_async_smem();                 // generic-proxy reads of the slot before its async-proxy (TMA) overwrite
                    mbar_arrive(bar_rempty + 8 * rs);
                }
                EPH(3);
                if (e_dbg & 16) { if (++ob == nob) ob = 0; continue; }
                // The store issued two chunks ago by this group read the buffer we are about to overwrite:
                // the leader waits until at most one of its bulk groups is still reading, then releases the group.
                // the store issued `nob` chunks ago read the buffer about to be overwritten
                if (LEAN) {
                    // waited above
                } else if (e_warp_store) {
                    if (elect_one()) { if (nob == 4) tma_store_wait_read<3>(); else tma_store_wait_read<1>(); }
                    __syncwarp();
                } else {
                    if (store_warp) { if (elect_one()) { if (nob == 4) tma_store_wait_read<3>(); else tma_store_wait_read<1>(); } __syncwarp(); }
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                }
                EPH(4);
                {
                    uint8_t* p = smem_raw + (out_base - smem_u32(smem_raw)) + (grp * nob + ob) * kEpiChunkBytes + row_off;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        uint4 u;
                        __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[c4 * 8 + 2 * e], v[c4 * 8 + 2 * e + 1]);
                        *reinterpret_cast<uint4*>(p + ((uint32_t(c4) ^ sw) << 4)) = u;
                    }
                }
                fence_proxy_async_smem();
                EPH(5);
                if (e_warp_store) {
                    __syncwarp();                                  // the slab is complete and visible to the async proxy
                    if (elect_one()) {
                        tma_store_4d(&tm_out, out_base + (grp * nob + ob) * kEpiChunkBytes + q * (32 * 64), out_c0 + j * kEpiCols,
                                     t.x0 + wx0, t.y0 + wy0, t.img + wz0);
                        tma_store_commit();
                    }
                    __syncwarp();
                } else {
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (store_warp) {
                        if (elect_one()) {
                            tma_store_4d(&tm_out, out_base + (grp * nob + ob) * kEpiChunkBytes, out_c0 + j * kEpiCols, t.x0, t.y0, t.img);
                            tma_store_commit();
                        }
                        __syncwarp();
                    }
                }
                EPH(6);
                if (++ob == nob) ob = 0;
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
#ifdef WIW_GEMM_PHASES
        if (eprof && lane == 0) { for (int i = 0; i < 8; ++i) g_gemm_prof[8 + i] = eph[i]; g_gemm_prof[16] = ech; g_gemm_prof[17] = eph[8]; g_gemm_prof[18] = eph[9]; }
#endif
        if (store_warp || e_warp_store) { if (elect_one()) tma_store_wait_all<0>(); __syncwarp(); }
        if (eprof && lane == 0) { g_gemm_prof[7] = clock64() - e_t0; g_gemm_prof[6] = e_wait; }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();      // the peer may still read this CTA's smem / signal its barriers
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2sm(tmem_base, g.tmem_cols); else
        tmem_dealloc(tmem_base, g.tmem_cols);
    }
}

// ----------------------------------------------------------------------------- host side
static int launch_conv_gemm(const wiw_gemm_desc* d, cudaStream_t stream) {
    WIW_REQUIRE(d != nullptr && d->a && d->w && d->out, "null pointer");
    WIW_REQUIRE(d->a_stride[0] == 1 && d->out_stride[0] == 1, "channel axis must be contiguous");
    const int bz = d->bz > 0 ? d->bz : 1;                 // images per M-tile (0 = 1: descriptors written before the field existed)
    WIW_REQUIRE(d->bx > 0 && d->by > 0 && d->bx * d->by * bz == kBlockM, "bx*by*bz must be 128");
    WIW_REQUIRE(d->bn >= 32 && d->bn <= 256 && d->bn % 32 == 0, "bn must be a multiple of 32 in [32,256]");
    WIW_REQUIRE(!d->geglu || d->bn % 64 == 0, "geglu needs bn % 64 == 0");
    WIW_REQUIRE(d->n_total % 32 == 0, "n_total must be a multiple of 32");
    WIW_REQUIRE(!d->geglu || d->n_total % d->bn == 0, "geglu needs n_total to be a multiple of bn");
    WIW_REQUIRE(d->num_taps >= 1 && d->num_taps <= 9 && d->chunks_per_tap >= 1, "bad tap spec");
    WIW_REQUIRE(d->k_total == (int64_t)d->num_taps * d->chunks_per_tap * kBlockK, "k_total != taps*chunks*64");
    const int64_t out_c = d->geglu ? d->n_total / 2 : d->n_total;
    WIW_REQUIRE(d->out_dim[0] == out_c, "out channel extent must equal n_total (n_total/2 for geglu)");
    WIW_REQUIRE(!d->tab || (d->tab_div > 0 && d->tab_mod > 0), "tab needs tab_div, tab_mod > 0");
    WIW_REQUIRE(!d->tab || !d->geglu, "tab with geglu is not supported");
    WIW_REQUIRE(!d->tab || (d->tab_stride % 8 == 0), "tab_stride must be a multiple of 8");
    WIW_REQUIRE(d->act >= 0 && d->act <= 3, "act must be 0 (none), 1 (SiLU), 2 (GELU) or 3 (quick GELU)");
    WIW_REQUIRE(!(d->act && d->geglu), "act with geglu is not supported");

    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.bx = d->bx; g.by = d->by; g.bz = bz;
    g.dim_x = (int)d->out_dim[1]; g.dim_y = (int)d->out_dim[2];
    g.tiles_x = (g.dim_x + g.bx - 1) / g.bx;
    g.tiles_y = (g.dim_y + g.by - 1) / g.by;
    g.n_tiles = (int)((d->n_total + d->bn - 1) / d->bn);          // a ragged last N-tile loads zero rows (TMA OOB fill) and skips its padded columns
    const long long m_tiles = (long long)g.tiles_x * g.tiles_y * ((d->out_dim[3] + bz - 1) / bz);
    WIW_REQUIRE(m_tiles * g.n_tiles < (1LL << 31), "too many tiles");
    // CTA pairs (cta_group::2, M = 256 per MMA, each CTA stages half of the weight tile) whenever the problem
    // fills the machine; small problems keep single-CTA tiles
    static const int force_cg = getenv("WIW_GEMM_CG") ? atoi(getenv("WIW_GEMM_CG")) : 0;      // development aid for A/B runs
    // Measured on B200 (profiles/r01_mma_cost.md, r01_gemm_pair_ab.md): a single-CTA tile is bound by shared-memory
    // bandwidth (TMA writes + UMMA reads of the same stage, 128 B/clk), so halving the weight bytes per CTA wins on
    // every shape except the tiny K=320, N=320 projection, where the pair handshakes cost ~3 %.
    const bool pair_shape = !(d->k_total <= 320 && d->n_total <= 320);
    int cg = (pair_shape && m_tiles * g.n_tiles >= 2LL * num_sms()) ? 2 : 1;
    if (force_cg == 1 || force_cg == 2) cg = force_cg;
    g.total_tiles = (int)(((m_tiles + cg - 1) / cg) * g.n_tiles);
    static const int dbg = getenv("WIW_GEMM_DBG") ? atoi(getenv("WIW_GEMM_DBG")) : 0;
    g.dbg = dbg;
    static const int epi = getenv("WIW_GEMM_EPI") ? atoi(getenv("WIW_GEMM_EPI")) : 1;       // development aid for A/B runs
    g.warp_store = epi ? 1 : 0;
    // one warp's 32 rows as a (pixel, row, image) box
    const int sbx = d->bx < 32 ? d->bx : 32, sby = (d->by < 32 / sbx) ? d->by : 32 / sbx, sbz = 32 / (sbx * sby);
    for (int q = 0; q < 4; ++q) {
        g.wx0[q] = (q * 32) % d->bx; g.wy0[q] = ((q * 32) / d->bx) % d->by; g.wz0[q] = (q * 32) / (d->bx * d->by);
    }
    g.bn = d->bn;
    g.num_taps = d->num_taps; g.chunks_per_tap = d->chunks_per_tap;
    g.geglu = d->geglu ? 1 : 0;
    g.out_cols_per_tile = g.geglu ? d->bn / 2 : d->bn;
    g.has_res = d->res ? 1 : 0; g.has_aux = d->aux ? 1 : 0;
    g.tab_div = d->tab_div; g.tab_mod = d->tab_mod; g.n_total = (int)d->n_total; g.out_total = (int)out_c;
    g.tab_stride = d->tab_stride > 0 ? d->tab_stride : (int)d->n_total; g.act = d->act;
    g.bias = d->bias; g.tab = static_cast<const __half*>(d->tab);
    g.s_acc = d->s_acc; g.s_res = d->s_res; g.s_aux = d->s_aux;
    for (int i = 0; i < d->num_taps; ++i) {
        g.tap_cbase[i] = d->tap_cbase[i]; g.tap_ox[i] = d->tap_ox[i]; g.tap_p[i] = d->tap_p[i]; g.tap_oy[i] = d->tap_oy[i];
    }
    int cols = 2 * ((g.bn <= 32) ? 32 : (g.bn <= 64 ? 64 : (g.bn <= 128 ? 128 : 256)));
    g.tmem_cols = cols;       // power of two in [64, 512]
    if (cg == 2) g.tmem_cols = 512;   // pair allocations always take the whole TMEM (as CUTLASS does): both CTAs must see the same base

    const int stage_bytes = kBlockM * kBlockK * 2 + (g.bn / cg) * kBlockK * 2;
    static const int obufs_env = getenv("WIW_GEMM_OBUFS") ? atoi(getenv("WIW_GEMM_OBUFS")) : 2;   // A/B on B200: 4 buffers per group = no gain (the epilogue is not waiting on store reads)
    g.out_bufs = (!g.has_res && !g.has_aux && obufs_env == 4) ? 4 : 2;
    // 8 residual slots (a whole tile's residual in flight) were tried and measured 5-17 % slower than 4: the shared memory is
    // worth more as A/B pipeline stages (profiles/r01_perf_gemm_v9_res_slots.txt).
    static const int rslots_env = getenv("WIW_GEMM_RSLOTS") ? atoi(getenv("WIW_GEMM_RSLOTS")) : 4;   // A/B on B200: 8 slots = 5-17 % SLOWER (they cost pipeline stages)
    g.res_slots = 4;
    {
        const int sb = kBlockM * kBlockK * 2 + (g.bn / cg) * kBlockK * 2;
        const int fixed8 = 2 * g.out_bufs * kEpiChunkBytes + 8 * kEpiChunkBytes + 512 + kBiasStageBytes;
        if (rslots_env == 8 && (g.has_res != g.has_aux) && (kSmemBudget - fixed8) / sb >= 4) g.res_slots = 8;
    }
    const int fixed = 2 * g.out_bufs * kEpiChunkBytes + (g.has_res ? g.res_slots * kEpiChunkBytes : 0) +
                      (g.has_aux ? g.res_slots * kEpiChunkBytes : 0) + 512 /*barriers*/ + kBiasStageBytes;
    int stages = (kSmemBudget - fixed) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    WIW_REQUIRE(stages >= 2, "not enough shared memory for 2 pipeline stages");
    g.stages = stages;
    const int smem_bytes = fixed + stages * stage_bytes;

    // ---- tensor maps
    CUtensorMap tm_a, tm_b, tm_out, tm_res, tm_aux;
    {
        uint64_t dims[5], strides[4];
        uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)d->bx, 1u, (uint32_t)d->by, (uint32_t)bz};
        for (int i = 0; i < 5; ++i) dims[i] = (uint64_t)d->a_dim[i];
        for (int i = 1; i < 5; ++i) strides[i - 1] = (uint64_t)d->a_stride[i] * 2;
        if (encode_tmap(&tm_a, d->a, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->k_total, (uint64_t)d->n_total};
        uint64_t strides[1] = {(uint64_t)d->k_total * 2};
        uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)(d->bn / cg)};
        if (encode_tmap(&tm_b, d->w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
    auto enc_out = [&](CUtensorMap* m, const void* p, const int64_t* st, bool warp_box) -> int {
        uint64_t dims[4], strides[3];
        uint32_t box[4] = {(uint32_t)kEpiCols, (uint32_t)(warp_box ? sbx : d->bx), (uint32_t)(warp_box ? sby : d->by), (uint32_t)(warp_box ? sbz : bz)};
        for (int i = 0; i < 4; ++i) dims[i] = (uint64_t)d->out_dim[i];
        for (int i = 1; i < 4; ++i) strides[i - 1] = (uint64_t)st[i] * 2;
        return encode_tmap(m, p, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    };
    if (enc_out(&tm_out, d->out, d->out_stride, g.warp_store != 0)) return 3;
    if (d->res) { if (enc_out(&tm_res, d->res, d->res_stride, false)) return 3; } else tm_res = tm_out;
    if (d->aux) { if (enc_out(&tm_aux, d->aux, d->aux_stride, false)) return 3; } else tm_aux = tm_out;

    static unsigned long long attr_set = 0;
    if (first_call_on_device(&attr_set)) {
        const void* all[8] = {(const void*)conv_gemm_kernel<1, 0, 0>, (const void*)conv_gemm_kernel<2, 0, 0>, (const void*)conv_gemm_kernel<1, 2, 0>,
                              (const void*)conv_gemm_kernel<2, 2, 0>, (const void*)conv_gemm_kernel<1, 0, 1>, (const void*)conv_gemm_kernel<2, 0, 1>,
                              (const void*)conv_gemm_kernel<1, 2, 1>, (const void*)conv_gemm_kernel<2, 2, 1>};
        for (const void* f : all) WIW_CHECK_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    }
    static const int lean_env = getenv("WIW_GEMM_LEAN") ? atoi(getenv("WIW_GEMM_LEAN")) : 1;    // development aid for A/B runs
    const bool lean = lean_env && d->act == 0 && !d->aux && (g.dbg & ~7) == 0 && g.warp_store && g.out_bufs == 2 && g.res_slots == 4;
    using kernel_t = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmArgs);
    kernel_t kern;
    if (cg == 1) kern = g.geglu ? (lean ? conv_gemm_kernel<1, 2, 1> : conv_gemm_kernel<1, 2, 0>) : (lean ? conv_gemm_kernel<1, 0, 1> : conv_gemm_kernel<1, 0, 0>);
    else         kern = g.geglu ? (lean ? conv_gemm_kernel<2, 2, 1> : conv_gemm_kernel<2, 2, 0>) : (lean ? conv_gemm_kernel<2, 0, 1> : conv_gemm_kernel<2, 0, 0>);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (cg == 1) {
        int grid = num_sms();
        if (grid > g.total_tiles) grid = g.total_tiles;
        cfg.gridDim = dim3(grid);
    } else {
        int pairs = num_sms() / 2;
        if (pairs > g.total_tiles) pairs = g.total_tiles;
        cfg.gridDim = dim3(2 * pairs);
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    WIW_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_a, tm_b, tm_out, tm_res, tm_aux, g));
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace wiw

extern "C" int wiw_debug_gemm_prof(long long* out24) {
    return cudaMemcpyFromSymbol(out24, wiw::g_gemm_prof, 24 * sizeof(long long)) == cudaSuccess ? 0 : 1;
}

extern "C" int wiw_conv_gemm(const wiw_gemm_desc* d, wiw_stream_t stream) {
    return wiw::launch_conv_gemm(d, static_cast<cudaStream_t>(stream));
}
