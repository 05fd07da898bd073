// Temporal self-attention over F <= 16 frames per (batch element, pixel, head), head_dim 64
// (TemporalBasicTransformerBlock.attn1, DP/models/attention.py:736 via AttnProcessor2_0): out = softmax(q k^T / 8) v.
//
// The problem per (b, s, h) is tiny (F x F scores, F = 14 in the deployed model), so the kernel is HBM-bound by
// design: it reads qkv [B][F][S][3C] once and writes out [B][F][S][C] once; the "(b f) s c <-> (b s) f c" permutes of the
// reference (attention.py:720-722,758-760) are pure index arithmetic.  One warp owns one problem at a time:
//   16-byte global loads of the F x 64 q / k / v rows -> padded smem -> ldmatrix fragments ->
//   S = Q K^T with 8 mma.sync.m16n8k16 (frames padded to 16) -> masked softmax in registers (quad shuffles) ->
//   O = P V with 8 mma.sync -> fp16 rows staged through smem -> 16-byte coalesced stores.
// (Legacy warp-level mma is the right tool for a 16x16x64 problem; tcgen05's 128-row tiles do not apply.)
#include "wiw_common.cuh"
#include "../../include/wiw.h"

namespace wiw {

constexpr int kTaWarps = 8;
constexpr int kTaRowH = 72;                       // 64 + 8 halves: 144-byte rows, conflict-free ldmatrix
constexpr int kTaTile = 16 * kTaRowH;             // one 16-row operand tile (halves)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(kTaWarps * 32)
temporal_attn_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int B, int F, int S, int C, float scale_log2) {
    extern __shared__ __align__(16) uint8_t smem_ta[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __half* sq = reinterpret_cast<__half*>(smem_ta) + (size_t)warp * 3 * kTaTile;
    __half* sk = sq + kTaTile;
    __half* sv = sk + kTaTile;
    // rows F..15 of every operand tile stay zero for the whole kernel
    for (int i = lane; i < 3 * kTaTile / 8; i += 32) reinterpret_cast<uint4*>(sq)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    const int heads = C / 64;
    const long long problems = (long long)B * S * heads;
    const long long row_stride = 3LL * C;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t sq_u = smem_u32(sq), sk_u = smem_u32(sk), sv_u = smem_u32(sv);
    // ldmatrix lane addresses (bytes): matrix m = lane/8, row lane%8
    const int lm = lane >> 3, lr = lane & 7;
    const uint32_t a_off = (uint32_t)(((lm & 1) * 8 + lr) * kTaRowH + (lm >> 1) * 8) * 2;       // A (Q): rows {0,8}, cols {0,8}
    const uint32_t bk_off = (uint32_t)(((lm >> 1) * 8 + lr) * kTaRowH + (lm & 1) * 8) * 2;      // B (K): n-tile lm/2, k-half lm&1
    const uint32_t bv_off = (uint32_t)(((lm & 1) * 8 + lr) * kTaRowH + (lm >> 1) * 8) * 2;      // B (V, transposed): k-half lm&1, n-tile lm/2
    const int nchunks = 3 * F * 8;

    for (long long p = (long long)blockIdx.x * kTaWarps + warp; p < problems; p += (long long)gridDim.x * kTaWarps) {
        const int h = (int)(p % heads);
        const long long bs = p / heads;
        const int s = (int)(bs % S), b = (int)(bs / S);
        const __half* src = qkv + ((long long)b * F * S + s) * row_stride + h * 64;
        // ---- global -> smem (q, k, v rows of F frames, 8 x 16 B each)
#pragma unroll 4
        for (int i = lane; i < nchunks; i += 32) {
            const int tsel = i / (F * 8), rem = i - tsel * (F * 8);
            const int f = rem >> 3, c = rem & 7;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (long long)f * S * row_stride + tsel * C + c * 8));
            *reinterpret_cast<uint4*>(sq + tsel * kTaTile + f * kTaRowH + c * 8) = u;
        }
        __syncwarp();
        // ---- S = Q K^T
        float sc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4], kb[4];
            ldsm_x4(sq_u + a_off + ks * 32, a[0], a[1], a[2], a[3]);
            ldsm_x4(sk_u + bk_off + ks * 32, kb[0], kb[1], kb[2], kb[3]);       // {nt0 k-lo, nt0 k-hi, nt1 k-lo, nt1 k-hi}
            mma_16816(sc[0], a, kb[0], kb[1]);
            mma_16816(sc[1], a, kb[2], kb[3]);
        }
        // ---- masked softmax over the F valid key frames (rows g and g+8; a row's 16 columns live in the 4 lanes of a quad)
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool ok = (nt * 8 + 2 * t + e) < F;
                sc[nt][e] = ok ? sc[nt][e] * scale_log2 : -INFINITY;
                sc[nt][2 + e] = ok ? sc[nt][2 + e] * scale_log2 : -INFINITY;
                m0 = fmaxf(m0, sc[nt][e]);
                m1 = fmaxf(m1, sc[nt][2 + e]);
            }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                sc[nt][e] = exp2f(sc[nt][e] - m0); l0 += sc[nt][e];
                sc[nt][2 + e] = exp2f(sc[nt][2 + e] - m1); l1 += sc[nt][2 + e];
            }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
        // C fragments of the two score tiles are exactly the A fragment of P (16 x 16)
        uint32_t pa[4] = {pack_h2(sc[0][0] * i0, sc[0][1] * i0), pack_h2(sc[0][2] * i1, sc[0][3] * i1),
                          pack_h2(sc[1][0] * i0, sc[1][1] * i0), pack_h2(sc[1][2] * i1, sc[1][3] * i1)};
        // ---- O = P V
        float o[8][4];
#pragma unroll
        for (int dn = 0; dn < 8; ++dn)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[dn][e] = 0.f;
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {                                        // two 8-wide d tiles per ldmatrix.x4.trans
            uint32_t vb[4];
            ldsm_x4_trans(sv_u + bv_off + dp * 32, vb[0], vb[1], vb[2], vb[3]);  // {d-tile 2dp: k-lo, k-hi, d-tile 2dp+1: k-lo, k-hi}
            mma_16816(o[2 * dp], pa, vb[0], vb[1]);
            mma_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
        }
        __syncwarp();                                                           // all lanes are done reading sq
        // ---- stage O (fp16) in the Q tile, then 16-byte coalesced row stores
#pragma unroll
        for (int dn = 0; dn < 8; ++dn) {
            *reinterpret_cast<uint32_t*>(sq + g * kTaRowH + dn * 8 + 2 * t) = pack_h2(o[dn][0], o[dn][1]);
            *reinterpret_cast<uint32_t*>(sq + (g + 8) * kTaRowH + dn * 8 + 2 * t) = pack_h2(o[dn][2], o[dn][3]);
        }
        __syncwarp();
        __half* dst = out + ((long long)b * F * S + s) * C + h * 64;
        for (int i = lane; i < F * 8; i += 32) {
            const int f = i >> 3, c = i & 7;
            *reinterpret_cast<uint4*>(dst + (long long)f * S * C + c * 8) = *reinterpret_cast<const uint4*>(sq + f * kTaRowH + c * 8);
        }
        __syncwarp();
        // rows F..15 of sq were overwritten by the staged O rows g+8 >= F: re-zero them for the next problem
        for (int i = lane; i < (16 - F) * 8; i += 32) {
            const int f = F + (i >> 3), c = i & 7;
            *reinterpret_cast<uint4*>(sq + f * kTaRowH + c * 8) = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
    }
}

}  // namespace wiw

using namespace wiw;

extern "C" int wiw_attention_temporal_mma(const void* qkv, void* out, int B, int F, int S, int C, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WIW_REQUIRE(qkv && out && C % 64 == 0 && F >= 1 && F <= 16, "bad args");
    const size_t smem = (size_t)kTaWarps * 3 * kTaTile * sizeof(__half);
    static unsigned long long attr_set = 0;
    if (first_call_on_device(&attr_set)) {
        WIW_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const long long problems = (long long)B * S * (C / 64);
    long long blocks = (problems + kTaWarps - 1) / kTaWarps;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    temporal_attn_mma_kernel<<<(unsigned)blocks, kTaWarps * 32, smem, stream>>>(
        static_cast<const __half*>(qkv), static_cast<__half*>(out), B, F, S, C, 0.125f * 1.4426950408889634f);
    WIW_CHECK_CUDA(cudaGetLastError());
    return 0;
}
