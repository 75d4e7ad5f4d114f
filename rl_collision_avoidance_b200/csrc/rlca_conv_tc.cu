// rlca_conv_tc.cu — CNNPolicy conv tower (model/net.py:21-22,42-44 of the reference) on the 5th-gen tensor cores.
//
//   conv1  Conv1d(3, 32, k5, s2, p1) + ReLU : 512 -> 255        conv2  Conv1d(32, 32, k3, s2, p1) + ReLU : 255 -> 128
//
// Both convolutions are dense contractions, so they run as tcgen05.mma.kind::tf32 with the 3xTF32 split
// (x = hi + lo, D += lo*hi + hi*lo + hi*hi; fp32 accumulation in TMEM) that keeps fp32 accuracy.
// One persistent CTA per SM walks over samples; per sample and for BOTH towers (actor, critic):
//
//   conv1   A1 = im2col of the scan, one 128-row tile: row j holds the 15 taps (+ a constant 1 that multiplies the
//           bias row of B) of the EVEN output position 2j in K columns 0..15 and of the ODD position 2j+1 in columns
//           16..31.  B1 = [64 = tower*32 + co][16] weights|bias.  Two M128 N64 K16 products -> D1even, D1odd in
//           TMEM: lane j = position, column = (tower, channel) - exactly the position-major, even/odd
//           de-interleaved layout conv2 wants as its A operand.
//   relu    worker warps pull D1 out of TMEM (tcgen05.ld), apply ReLU, split hi/lo and store the E (even) and
//           O (odd) tiles [128 positions x 32 channels] of each tower as 128B-swizzled K-major smem tiles.
//   conv2   out[q] = W_k1 h1[2q] + W_k2 h1[2q+1] + W_k0 h1[2q-1]:   Da = O.[W_k2|W_k0] (N64) then Da[:, :32] += E.W_k1.
//           The k0 tap is produced one row too low (row q holds W_k0 h1[2q+1]); the epilogue adds row q-1 with one
//           warp shuffle instead of building a shifted copy of the operand in shared memory.
//   store   relu(Da + shift(Db) + b2) -> F (flatten order c*128+q) and its tf32 hi/lo split for the fc1 GEMM.
//
// Warp roles: warps 0-7 workers (im2col build / TMEM drain; warps 0-3 tower 0, 4-7 tower 1, TMEM lane quadrant =
// warp % 4), warp 8 lane 0 issues every MMA.  mbarriers order workers <-> tensor pipe; the issue order
// conv2(n, t0), conv2(n, t1), conv1(n+1) lets the workers' epilogue / next im2col overlap the tensor work.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rlca.h"
#include "rlca_common.cuh"
#include "rlca_conv_tc.cuh"
#include "rlca_tc_ptx.cuh"

using namespace rlca_ptx;

namespace {

constexpr int FEAT = 4096;
constexpr int TILE_F = 128 * 32;                  // floats in one [128 x 32] operand tile (16 KB)
// ---- weight image (global, prepared once per weight change; float offsets).  Every tile is already 128B-swizzled.
constexpr int W_B1 = 0;                           // [64 rows = tower*32+co][32]: cols 0..15 hi(w1|b1), 16..31 lo
constexpr int W_T0 = 64 * 32;                     // per tower: B2O_hi[64][32] B2O_lo B2E_hi[32][32] B2E_lo
constexpr int W_TOWER = 2 * 64 * 32 + 2 * 32 * 32;
constexpr int W_BIAS2 = W_T0 + 2 * W_TOWER;       // [2][32] conv2 bias
constexpr int W_FLOATS = W_BIAS2 + 64;            // 14400 floats = 57600 B
// ---- shared memory (byte offsets from a 1024-aligned base)
constexpr int OFF_W = 0;
constexpr int OFF_A1 = 57 * 1024;                 // A1 hi, lo
constexpr int OFF_A2 = OFF_A1 + 2 * TILE_F * 4;   // [tower][E_hi, E_lo, O_hi, O_lo]
constexpr int XS_PITCH = 520;                     // padded scan row: xs[c][4 + i] = x[c][i], zeros around
constexpr int OFF_XS = OFF_A2 + 8 * TILE_F * 4;
constexpr int OFF_XCHG = OFF_XS + 3 * XS_PITCH * 4;
constexpr int OFF_BAR = OFF_XCHG + 2 * 4 * 32 * 4;
constexpr int SMEM_USED = OFF_BAR + 128;
constexpr size_t SMEM_BYTES = SMEM_USED + 1024;   // + alignment slack
static_assert(W_FLOATS * 4 <= OFF_A1, "weight image overlaps A1");
static_assert(SMEM_BYTES <= 227 * 1024, "conv tc kernel exceeds the 227 KB shared-memory limit");

constexpr int NWORK = 256;
constexpr int NTHREADS = NWORK + 32;
constexpr uint32_t TMEM_COLS = 256;               // D1: cols 0..127 (even|odd x 2 towers x 32), D2: 128 + t*64 (+32: k0 tap)

enum { BAR_A1 = 0, BAR_A2 = 1 /* +t */, BAR_C1 = 3, BAR_C2 = 4 /* +t */, BAR_X = 6, BAR_W = 7 };

struct ConvTcWeights {
    const float *cv1w[2], *cv1b[2], *cv2w[2], *cv2b[2];
};

__global__ void conv_tc_prep_kernel(ConvTcWeights w, float *__restrict__ img)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W_FLOATS) return;
    float val;
    int dst;
    bool lo;
    // (tower pointers selected with ?: - indexing a kernel-parameter array with a runtime value forces a local copy)
    if (i < W_T0) {
        const int n = i >> 5, c = i & 31, t = n >> 5, co = n & 31, k = c & 15;
        lo = c >= 16;
        val = k < 15 ? (t ? w.cv1w[1] : w.cv1w[0])[co * 15 + k] : (t ? w.cv1b[1] : w.cv1b[0])[co];
        dst = W_B1 + sw128_index(n, c);
    } else if (i < W_BIAS2) {
        const int r = i - W_T0, t = r / W_TOWER, q = r - t * W_TOWER;
        if (q < 2 * 64 * 32) {            // O-pass operand: rows 0..31 tap k=2 (-> Da), rows 32..63 tap k=0 (-> Db)
            lo = q >= 64 * 32;
            const int e = q & (64 * 32 - 1), n = e >> 5, ci = e & 31, co = n & 31, tap = n < 32 ? 2 : 0;
            val = (t ? w.cv2w[1] : w.cv2w[0])[co * 96 + ci * 3 + tap];
            dst = W_T0 + t * W_TOWER + (lo ? 64 * 32 : 0) + sw128_index(n, ci);
        } else {                          // E-pass operand: tap k=1
            const int q2 = q - 2 * 64 * 32;
            lo = q2 >= 32 * 32;
            const int e = q2 & (32 * 32 - 1), co = e >> 5, ci = e & 31;
            val = (t ? w.cv2w[1] : w.cv2w[0])[co * 96 + ci * 3 + 1];
            dst = W_T0 + t * W_TOWER + 2 * 64 * 32 + (lo ? 32 * 32 : 0) + sw128_index(co, ci);
        }
    } else {
        const int r = i - W_BIAS2;
        img[i] = ((r >> 5) ? w.cv2b[1] : w.cv2b[0])[r & 31];
        return;
    }
    const float h = tf32_hi(val);
    img[dst] = lo ? val - h : h;
}

// im2col row (half H: 0 = even output position 2j, 1 = odd position 2j+1) -> 4 chunks of the A1 hi / lo tiles
template <int H>
__device__ __forceinline__ void build_a1_row(const float *__restrict__ xs, float *__restrict__ a1hi, float *__restrict__ a1lo,
                                             int j)
{
    float v[16];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const float4 p = *reinterpret_cast<const float4 *>(xs + ci * XS_PITCH + 4 * j);
        const float4 q = *reinterpret_cast<const float4 *>(xs + ci * XS_PITCH + 4 * j + 4);
        const float4 r = *reinterpret_cast<const float4 *>(xs + ci * XS_PITCH + 4 * j + 8);
        const float w[12] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w, r.x, r.y, r.z, r.w};
        // padded index of tap kk: (2p + kk - 1) + 4 with p = 2j + H  ->  4j + 3 + 2H + kk
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) v[ci * 5 + kk] = w[3 + 2 * H + kk];
    }
    v[15] = 1.0f;                      // multiplies the bias row of B1
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int idx = j * 32 + (((H * 4 + c) ^ (j & 7)) << 2);
        float4 hi, lo;
        hi.x = tf32_hi(v[4 * c + 0]); hi.y = tf32_hi(v[4 * c + 1]); hi.z = tf32_hi(v[4 * c + 2]); hi.w = tf32_hi(v[4 * c + 3]);
        lo.x = v[4 * c + 0] - hi.x; lo.y = v[4 * c + 1] - hi.y; lo.z = v[4 * c + 2] - hi.z; lo.w = v[4 * c + 3] - hi.w;
        *reinterpret_cast<float4 *>(a1hi + idx) = hi;
        *reinterpret_cast<float4 *>(a1lo + idx) = lo;
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv_tower_fwd_tc_kernel(const float *__restrict__ obs, const float *__restrict__ img, float *__restrict__ F,
                         float *__restrict__ Fs, int nb)
{
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the swizzle atoms, by pointer arithmetic so the compiler keeps the shared address space
    uint8_t *sm = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    float *wimg = reinterpret_cast<float *>(sm + OFF_W);
    float *a1hi = reinterpret_cast<float *>(sm + OFF_A1), *a1lo = a1hi + TILE_F;
    float *a2 = reinterpret_cast<float *>(sm + OFF_A2);
    float *xs = reinterpret_cast<float *>(sm + OFF_XS);
    float *xchg = reinterpret_cast<float *>(sm + OFF_XCHG);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + OFF_BAR);
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bars[BAR_A1], NWORK);
        mbar_init(&bars[BAR_A2 + 0], NWORK / 2);
        mbar_init(&bars[BAR_A2 + 1], NWORK / 2);
        mbar_init(&bars[BAR_C1], 1);
        mbar_init(&bars[BAR_C2 + 0], 1);
        mbar_init(&bars[BAR_C2 + 1], 1);
        mbar_init(&bars[BAR_X], 1);
        mbar_init(&bars[BAR_W], 1);
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(tmem_ptr, TMEM_COLS);
    if (tid < 3 * 8) {                  // zero padding of the staged scan: 4 floats before and after every channel
        const int c = tid >> 3, k = tid & 7;
        xs[c * XS_PITCH + (k < 4 ? k : 512 + k)] = 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const int first = blockIdx.x, stride = gridDim.x;
    // the scan of sample n, three 2 KB rows, lands in xs through the bulk-copy engine (one thread issues)
    auto load_x = [&](int n) {
        mbar_expect_tx(&bars[BAR_X], 3 * 2048);
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_bulk_g2s(xs + c * XS_PITCH + 4, obs + (size_t)n * 1536 + c * 512, 2048, &bars[BAR_X]);
    };

    if (warp == 8) {
        // =================================================================== MMA issuer
        if (lane == 0) {
            const uint32_t id64 = umma_idesc_tf32(128, 64), id32 = umma_idesc_tf32(128, 32);
            const uint32_t sA1h = smem_addr(a1hi), sA1l = smem_addr(a1lo), sW = smem_addr(wimg);
            auto issue_conv1 = [&]() {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t d = tmem + (uint32_t)h * 64;
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint32_t ao = (uint32_t)h * 64 + kb * 32, bh = sW + W_B1 * 4 + kb * 32, bl = bh + 64;
                        umma_tf32(d, umma_desc_sw128(sA1l + ao), umma_desc_sw128(bh), id64, kb ? 1u : 0u);
                        umma_tf32(d, umma_desc_sw128(sA1h + ao), umma_desc_sw128(bl), id64, 1u);
                        umma_tf32(d, umma_desc_sw128(sA1h + ao), umma_desc_sw128(bh), id64, 1u);
                    }
                }
                umma_commit(&bars[BAR_C1]);
            };
            auto issue_conv2 = [&](int t) {
                const uint32_t sEh = smem_addr(a2 + (size_t)(t * 4 + 0) * TILE_F), sEl = sEh + TILE_F * 4;
                const uint32_t sOh = sEh + 2 * TILE_F * 4, sOl = sEh + 3 * TILE_F * 4;
                const uint32_t bOh = sW + (W_T0 + t * W_TOWER) * 4, bOl = bOh + 64 * 32 * 4;
                const uint32_t bEh = bOh + 2 * 64 * 32 * 4, bEl = bEh + 32 * 32 * 4;
                const uint32_t d = tmem + 128 + (uint32_t)t * 64;
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {         // O tile: taps k=2 (cols 0..31) and k=0 (cols 32..63)
                    const uint32_t o = kb * 32;
                    umma_tf32(d, umma_desc_sw128(sOl + o), umma_desc_sw128(bOh + o), id64, kb ? 1u : 0u);
                    umma_tf32(d, umma_desc_sw128(sOh + o), umma_desc_sw128(bOl + o), id64, 1u);
                    umma_tf32(d, umma_desc_sw128(sOh + o), umma_desc_sw128(bOh + o), id64, 1u);
                }
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {         // E tile: tap k=1, accumulated onto cols 0..31
                    const uint32_t o = kb * 32;
                    umma_tf32(d, umma_desc_sw128(sEl + o), umma_desc_sw128(bEh + o), id32, 1u);
                    umma_tf32(d, umma_desc_sw128(sEh + o), umma_desc_sw128(bEl + o), id32, 1u);
                    umma_tf32(d, umma_desc_sw128(sEh + o), umma_desc_sw128(bEh + o), id32, 1u);
                }
                umma_commit(&bars[BAR_C2 + t]);
            };
            int it = 0;
            mbar_expect_tx(&bars[BAR_W], W_FLOATS * 4);
            tma_bulk_g2s(wimg, img, W_FLOATS * 4, &bars[BAR_W]);
            if (first < nb) {
                load_x(first);
                mbar_wait(&bars[BAR_W], 0);
                mbar_wait(&bars[BAR_A1], 0);
                tc_fence_after();
                if (first + stride < nb) load_x(first + stride);      // every worker has consumed xs
                issue_conv1();
            }
            for (int n = first; n < nb; n += stride, ++it) {
                const uint32_t ph = (uint32_t)it & 1u;
                for (int t = 0; t < 2; ++t) {
                    mbar_wait(&bars[BAR_A2 + t], ph);
                    tc_fence_after();
                    issue_conv2(t);
                }
                if (n + stride < nb) {
                    mbar_wait(&bars[BAR_A1], ph ^ 1u);
                    tc_fence_after();
                    if (n + 2 * stride < nb) load_x(n + 2 * stride);
                    issue_conv1();
                }
            }
        }
    } else {
        // =================================================================== workers
        const int lq = warp & 3, grp = warp >> 2;        // grp = im2col half while building A1, tower afterwards
        const int j = lq * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(lq * 32) << 16);
        const float *bias2 = wimg + W_BIAS2 + grp * 32;
        auto build_a1 = [&](uint32_t xph) {
            mbar_wait(&bars[BAR_X], xph);
            if (grp == 0) build_a1_row<0>(xs, a1hi, a1lo, j);
            else build_a1_row<1>(xs, a1hi, a1lo, j);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars[BAR_A1]);
        };
        int it = 0;
        if (first < nb) build_a1(0);
        for (int n = first; n < nb; n += stride, ++it) {
            const uint32_t ph = (uint32_t)it & 1u;
            // ---- conv1 result -> ReLU -> E / O operand tiles of this warp group's tower
            mbar_wait(&bars[BAR_C1], ph);
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld32(lane_base + (uint32_t)(h * 64 + grp * 32), r);
                tmem_ld_wait();
                float *th = a2 + (size_t)(grp * 4 + h * 2) * TILE_F, *tl = th + TILE_F;
                const bool pad = (h == 1 && j == 127);      // position 255 is conv2's right zero padding
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 x;
                    x.x = pad ? 0.f : fmaxf(__uint_as_float(r[4 * c + 0]), 0.f);
                    x.y = pad ? 0.f : fmaxf(__uint_as_float(r[4 * c + 1]), 0.f);
                    x.z = pad ? 0.f : fmaxf(__uint_as_float(r[4 * c + 2]), 0.f);
                    x.w = pad ? 0.f : fmaxf(__uint_as_float(r[4 * c + 3]), 0.f);
                    const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                    const int idx = j * 32 + ((c ^ (j & 7)) << 2);
                    *reinterpret_cast<float4 *>(th + idx) = hi;
                    *reinterpret_cast<float4 *>(tl + idx) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars[BAR_A2 + grp]);
            // ---- next sample's im2col while the tensor pipe runs conv2 (conv1(n) has retired: A1 and xs are free)
            if (n + stride < nb) build_a1(ph ^ 1u);
            // ---- conv2 result: out[q] = Da[q] + Db[q-1] + b2, ReLU, store F and its tf32 split
            if (it == 0) mbar_wait(&bars[BAR_W], 0);          // bias2 comes from the weight image
            mbar_wait(&bars[BAR_C2 + grp], ph);
            tc_fence_after();
            uint32_t da[32], db[32];
            tmem_ld32(lane_base + (uint32_t)(128 + grp * 64), da);
            tmem_ld32(lane_base + (uint32_t)(128 + grp * 64 + 32), db);
            tmem_ld_wait();
            float *xc = xchg + (grp * 4 + lq) * 32;
            if (lane == 31) {
#pragma unroll
                for (int co = 0; co < 32; ++co) xc[co] = __uint_as_float(db[co]);
            }
            named_bar_sync(2 + grp, NWORK / 2);
            float *out = F + ((size_t)grp * nb + n) * FEAT + j;
            float *out_hi = Fs ? Fs + ((size_t)(2 * grp) * nb + n) * FEAT + j : nullptr;
            float *out_lo = Fs ? Fs + ((size_t)(2 * grp + 1) * nb + n) * FEAT + j : nullptr;
#pragma unroll
            for (int co = 0; co < 32; ++co) {
                float up = __uint_as_float(__shfl_up_sync(0xffffffffu, db[co], 1));
                if (lane == 0) up = lq ? xc[co - 32] : 0.0f;      // row q-1 lives in the previous warp (q = 0: left padding)
                const float v = fmaxf(__uint_as_float(da[co]) + up + bias2[co], 0.0f);
                out[co * 128] = v;
                if (Fs) {
                    const float hi = tf32_hi(v);
                    out_hi[co * 128] = hi;
                    out_lo[co * 128] = v - hi;
                }
            }
            // the exchange buffer is rewritten only after every thread of the group passed the next BAR_A2 arrival
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, TMEM_COLS);
    }
}


// =====================================================================================================================
// Backward of the conv tower on the tensor cores.  One persistent CTA per (SM, tower); per sample:
//
//   conv1   recomputed exactly as in the forward kernel (N = 32: this CTA's tower only) -> h1 (even | odd) in TMEM
//   P1      dh1 "pre-scatter":  Dp[q][(tap, c1)] = sum_co g2[q][co] W2[co][c1][tap]         A = G  [128 q  x 32 co]
//   P2      dW2 += sum_q HT[(tap, c1)][q] g2T[co][q]    (K = positions)                       A = HT [96 x 128 q]
//   P3      dW1 | db1 += sum_j g1T[(half, c1)][j] im2colT[(half, k)][j]   (ones column -> bias gradient)
//
// tcgen05.mma.kind::tf32 only takes K-major operands (an MN-major instruction descriptor is silently a no-op on
// sm_100a: tools/probes/probe_mn_major.cu), so the operands whose contraction index is the position are written
// TRANSPOSED by the workers: a warp holds 32 consecutive positions, so one transposed row segment is a conflict-free
// 128-byte store.  dW2 / dW1 accumulate over all samples of the CTA in TMEM and leave once, as one partial per CTA
// (conv_part_reduce_kernel then sums <= 74 partials per tower instead of one per sample).  g2 = dF * (F > 0) is
// formed on load; g1 = dh1 * (h1 > 0) uses mask bits kept in a register from the conv1 drain.
namespace bw {
constexpr int WB_B1 = 0;                          // [32 co][32]: cols 0..15 hi(w1|b1), 16..31 lo
constexpr int WB_W2H = 32 * 32;                   // [96 = blk*32 + c1][32 co] hi, blk 0/1/2 = tap 1/2/0
constexpr int WB_W2L = WB_W2H + 96 * 32;
constexpr int WB_TOWER = WB_W2L + 96 * 32;        // 7168 floats = 28 KB per tower
constexpr int OFF_W = 0;
constexpr int OFF_A1 = WB_TOWER * 4;              // A1 hi | lo; later the same bytes hold im2colT hi | lo
constexpr int OFF_HT = OFF_A1 + 2 * TILE_F * 4;   // HT hi | lo (4 k-atoms of 96 rows each); later g1T hi | lo
constexpr int HT_ATOM = 96 * 32, HT_LO = 4 * HT_ATOM;
constexpr int G1_ATOM = 64 * 32, G1_LO = 4 * G1_ATOM;
constexpr int IM_ATOM = 32 * 32, IM_LO = 4 * IM_ATOM;
constexpr int OFF_GT = OFF_HT + 2 * HT_LO * 4;    // g2T hi | lo (4 k-atoms of 32 rows)
constexpr int GT_ATOM = 32 * 32, GT_LO = 4 * GT_ATOM;
constexpr int OFF_G = OFF_GT + 2 * GT_LO * 4;     // G hi | lo  [128 q x 32 co]
constexpr int OFF_XCHG = OFF_G + 2 * TILE_F * 4;
constexpr int OFF_BAR = OFF_XCHG + 4 * 32 * 4;
constexpr int SMEM_USED = OFF_BAR + 128;
constexpr size_t SMEM_BYTES = SMEM_USED + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "conv tc backward exceeds the 227 KB shared-memory limit");
static_assert(OFF_A1 % 1024 == 0 && OFF_HT % 1024 == 0 && OFF_GT % 1024 == 0 && OFF_G % 1024 == 0, "tile alignment");
constexpr int PART = 3616;                        // cv2w 3072 | cv2b 32 | cv1w 480 | cv1b 32  (CONV_PART of rlca_policy.cu)
enum { B_A1 = 0, B_G, B_G1, B_C1, B_P1, B_P2, B_P3, B_W };
constexpr uint32_t TM_D1 = 0, TM_DP = 64, TM_ACC2 = 160, TM_ACC1 = 192, TM_COLS = 256;
}  // namespace bw

__global__ void conv_tc_bwd_prep_kernel(ConvTcWeights w, float *__restrict__ img)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (i >= bw::WB_TOWER) return;
    float val;
    int dst;
    bool lo;
    if (i < bw::WB_W2H) {
        const int co = i >> 5, c = i & 31, k = c & 15;
        lo = c >= 16;
        val = k < 15 ? (t ? w.cv1w[1] : w.cv1w[0])[co * 15 + k] : (t ? w.cv1b[1] : w.cv1b[0])[co];
        dst = bw::WB_B1 + sw128_index(co, c);
    } else {
        const int r = i - bw::WB_W2H;
        lo = r >= 96 * 32;
        const int e = lo ? r - 96 * 32 : r, n = e >> 5, co = e & 31, blk = n >> 5, c1 = n & 31;
        const int tap = blk == 0 ? 1 : (blk == 1 ? 2 : 0);
        val = (t ? w.cv2w[1] : w.cv2w[0])[co * 96 + c1 * 3 + tap];
        dst = (lo ? bw::WB_W2L : bw::WB_W2H) + sw128_index(n, co);
    }
    const float h = tf32_hi(val);
    img[(size_t)t * bw::WB_TOWER + dst] = lo ? val - h : h;
}

// the 16 im2col values (15 taps + the bias one) of output position 2j + H, straight from global memory
template <int H>
__device__ __forceinline__ void load_im2col(const float *__restrict__ x, int j, float (&v)[16])
{
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const float *row = x + ci * 512;
        const float4 b = __ldg(reinterpret_cast<const float4 *>(row) + j);
        if (H == 0) {
            const float a = j > 0 ? __ldg(row + 4 * j - 1) : 0.0f;          // left zero padding of Conv1d
            v[ci * 5 + 0] = a; v[ci * 5 + 1] = b.x; v[ci * 5 + 2] = b.y; v[ci * 5 + 3] = b.z; v[ci * 5 + 4] = b.w;
        } else {
            const float4 c = j < 127 ? __ldg(reinterpret_cast<const float4 *>(row) + j + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[ci * 5 + 0] = b.y; v[ci * 5 + 1] = b.z; v[ci * 5 + 2] = b.w; v[ci * 5 + 3] = c.x; v[ci * 5 + 4] = c.y;
        }
    }
    v[15] = 1.0f;
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv_tower_bwd_tc_kernel(const float *__restrict__ obs, const float *__restrict__ img, const float *__restrict__ dF,
                         const float *__restrict__ Fm, float *__restrict__ part, int nb, int slots)
{
    using namespace bw;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    float *wimg = reinterpret_cast<float *>(sm + bw::OFF_W);
    float *a1 = reinterpret_cast<float *>(sm + bw::OFF_A1);      // hi, lo at + TILE_F  (im2colT: hi, lo at + IM_LO)
    float *ht = reinterpret_cast<float *>(sm + OFF_HT);
    float *gt = reinterpret_cast<float *>(sm + OFF_GT);
    float *gq = reinterpret_cast<float *>(sm + OFF_G);       // hi, lo at + TILE_F
    float *xchg = reinterpret_cast<float *>(sm + bw::OFF_XCHG);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + bw::OFF_BAR);
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = blockIdx.x >= slots ? 1 : 0, slot = blockIdx.x - t * slots;
    if (tid == 0) {
        mbar_init(&bars[B_A1], NWORK); mbar_init(&bars[B_G], NWORK); mbar_init(&bars[B_G1], NWORK);
        mbar_init(&bars[B_C1], 1); mbar_init(&bars[B_P1], 1); mbar_init(&bars[B_P2], 1); mbar_init(&bars[B_P3], 1);
        mbar_init(&bars[B_W], 1);
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(tmem_ptr, TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const int first = slot, stride = slots;

    if (warp == 8) {
        // =================================================================== MMA issuer
        if (lane == 0) {
            const uint32_t id32 = umma_idesc_tf32(128, 32), id96 = umma_idesc_tf32(128, 96);
            const uint32_t sW = smem_addr(wimg), sA1 = smem_addr(a1), sHT = smem_addr(ht), sGT = smem_addr(gt), sG = smem_addr(gq);
            mbar_expect_tx(&bars[B_W], WB_TOWER * 4);
            tma_bulk_g2s(wimg, img + (size_t)t * WB_TOWER, WB_TOWER * 4, &bars[B_W]);
            mbar_wait(&bars[B_W], 0);
            int it = 0;
            for (int n = first; n < nb; n += stride, ++it) {
                const uint32_t ph = (uint32_t)it & 1u;
                // ---- conv1 (this tower): D1 even -> cols 0..31, odd -> 32..63
                mbar_wait(&bars[B_A1], ph);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint32_t ao = (uint32_t)h * 64 + kb * 32, bh = sW + WB_B1 * 4 + kb * 32, bl = bh + 64;
                        const uint32_t d = tmem + TM_D1 + (uint32_t)h * 32;
                        umma_tf32(d, umma_desc_sw128(sA1 + TILE_F * 4 + ao), umma_desc_sw128(bh), id32, kb ? 1u : 0u);
                        umma_tf32(d, umma_desc_sw128(sA1 + ao), umma_desc_sw128(bl), id32, 1u);
                        umma_tf32(d, umma_desc_sw128(sA1 + ao), umma_desc_sw128(bh), id32, 1u);
                    }
                umma_commit(&bars[B_C1]);
                // ---- P1: Dp[q][(blk, c1)] = G . W2r^T  (K = co = 32)
                mbar_wait(&bars[B_G], ph);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    const uint32_t o = kb * 32, d = tmem + TM_DP;
                    umma_tf32(d, umma_desc_sw128(sG + TILE_F * 4 + o), umma_desc_sw128(sW + WB_W2H * 4 + o), id96, kb ? 1u : 0u);
                    umma_tf32(d, umma_desc_sw128(sG + o), umma_desc_sw128(sW + WB_W2L * 4 + o), id96, 1u);
                    umma_tf32(d, umma_desc_sw128(sG + o), umma_desc_sw128(sW + WB_W2H * 4 + o), id96, 1u);
                }
                umma_commit(&bars[B_P1]);
                // ---- P2: ACC2[(blk, c1)][co] += HT . g2T^T  (K = q = 128: 4 k-atoms x 4 k-blocks)
#pragma unroll 1
                for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint32_t ao = (uint32_t)ka * HT_ATOM * 4 + kb * 32, bo = (uint32_t)ka * GT_ATOM * 4 + kb * 32;
                        const uint32_t d = tmem + TM_ACC2;
                        umma_tf32(d, umma_desc_sw128(sHT + HT_LO * 4 + ao), umma_desc_sw128(sGT + bo), id32, (it | ka | kb) ? 1u : 0u);
                        umma_tf32(d, umma_desc_sw128(sHT + ao), umma_desc_sw128(sGT + GT_LO * 4 + bo), id32, 1u);
                        umma_tf32(d, umma_desc_sw128(sHT + ao), umma_desc_sw128(sGT + bo), id32, 1u);
                    }
                umma_commit(&bars[B_P2]);
                // ---- P3: ACC1[(half, c1)][(half', k)] += g1T . im2colT^T  (K = j = 128)
                mbar_wait(&bars[B_G1], ph);
                tc_fence_after();
#pragma unroll 1
                for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint32_t ao = (uint32_t)ka * G1_ATOM * 4 + kb * 32, bo = (uint32_t)ka * IM_ATOM * 4 + kb * 32;
                        const uint32_t d = tmem + TM_ACC1;
                        umma_tf32(d, umma_desc_sw128(sHT + G1_LO * 4 + ao), umma_desc_sw128(sA1 + bo), id32, (it | ka | kb) ? 1u : 0u);
                        umma_tf32(d, umma_desc_sw128(sHT + ao), umma_desc_sw128(sA1 + IM_LO * 4 + bo), id32, 1u);
                        umma_tf32(d, umma_desc_sw128(sHT + ao), umma_desc_sw128(sA1 + bo), id32, 1u);
                    }
                umma_commit(&bars[B_P3]);
            }
        }
    } else {
        // =================================================================== workers
        const int lq = warp & 3, half = warp >> 2;       // half: im2col half / even-odd h1 / channel half of g2
        const int j = lq * 32 + lane, ka = lq, jc = lane;   // position j = k-atom lq, column `lane` of a transposed tile
        const uint32_t lane_base = tmem + ((uint32_t)(lq * 32) << 16);
        float acc_b2[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc_b2[c] = 0.0f;
        int it = 0;
        for (int n = first; n < nb; n += stride, ++it) {
            const uint32_t ph = (uint32_t)it & 1u;
            // ---- im2col values of this thread's output position (kept in registers until im2colT is written)
            float v[16];
            if (half == 0) load_im2col<0>(obs + (size_t)n * 1536, j, v);
            else load_im2col<1>(obs + (size_t)n * 1536, j, v);
            // ---- g2 = dF * (F > 0) for q = j, channels half*16 ..: G row (K-major in co) and g2T (K-major in q)
            {
                const size_t base = ((size_t)t * nb + n) * FEAT + (size_t)(half * 16) * 128 + j;
                float g[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float f = __ldg(Fm + base + c * 128), d = __ldg(dF + base + c * 128);
                    g[c] = f > 0.0f ? d : 0.0f;
                    acc_b2[c] += g[c];
                }
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 x = make_float4(g[4 * c4], g[4 * c4 + 1], g[4 * c4 + 2], g[4 * c4 + 3]);
                    const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                    const int idx = j * 32 + (((half * 4 + c4) ^ (j & 7)) << 2);
                    *reinterpret_cast<float4 *>(gq + idx) = hi;
                    *reinterpret_cast<float4 *>(gq + TILE_F + idx) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int idx = ka * GT_ATOM + sw128_index(half * 16 + c, jc);
                    const float hi = tf32_hi(g[c]);
                    gt[idx] = hi;
                    gt[GT_LO + idx] = g[c] - hi;
                }
            }
            // ---- A1 (aliases the previous sample's im2colT: wait for its P3)
            if (it > 0) mbar_wait(&bars[B_P3], ph ^ 1u);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int idx = j * 32 + (((half * 4 + c) ^ (j & 7)) << 2);
                const float4 hi = make_float4(tf32_hi(v[4 * c]), tf32_hi(v[4 * c + 1]), tf32_hi(v[4 * c + 2]), tf32_hi(v[4 * c + 3]));
                *reinterpret_cast<float4 *>(a1 + idx) = hi;
                *reinterpret_cast<float4 *>(a1 + TILE_F + idx) =
                    make_float4(v[4 * c] - hi.x, v[4 * c + 1] - hi.y, v[4 * c + 2] - hi.z, v[4 * c + 3] - hi.w);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars[B_A1]);
            // ---- h1 = relu(conv1): transposed into HT (E rows 0..31 | O rows 32..63 | O shifted by one q, rows 64..95)
            mbar_wait(&bars[B_C1], ph);
            tc_fence_after();
            uint32_t hmask = 0u;
            {
                uint32_t r[32];
                tmem_ld32(lane_base + TM_D1 + (uint32_t)half * 32, r);
                tmem_ld_wait();
                const bool pad = (half == 1 && j == 127);           // position 255 = conv2's right zero padding
#pragma unroll
                for (int c1 = 0; c1 < 32; ++c1) {
                    const float x = pad ? 0.0f : fmaxf(__uint_as_float(r[c1]), 0.0f);
                    hmask |= (x > 0.0f ? 1u : 0u) << c1;
                    const float hi = tf32_hi(x), lo = x - hi;
                    const int idx = ka * HT_ATOM + sw128_index(half * 32 + c1, jc);
                    ht[idx] = hi;
                    ht[HT_LO + idx] = lo;
                    if (half == 1) {
                        // row 64 + c1 holds O[q - 1]: this thread's value belongs to column j + 1; thread 127 writes the
                        // q = 0 column (left padding) instead
                        const int q = j == 127 ? 0 : j + 1;
                        const int idx2 = (q >> 5) * HT_ATOM + sw128_index(64 + c1, q & 31);
                        ht[idx2] = j == 127 ? 0.0f : hi;
                        ht[HT_LO + idx2] = j == 127 ? 0.0f : lo;
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars[B_G]);
            // ---- dh1 from P1, masked by relu(conv1) -> g1 (even positions: half 0, odd: half 1)
            mbar_wait(&bars[B_P1], ph);
            tc_fence_after();
            float g1[32];
            if (half == 0) {
                uint32_t r[32];
                tmem_ld32(lane_base + TM_DP, r);                    // tap 1: dh1[2j] = g2[j] . W_k1
                tmem_ld_wait();
#pragma unroll
                for (int c1 = 0; c1 < 32; ++c1) g1[c1] = ((hmask >> c1) & 1u) ? __uint_as_float(r[c1]) : 0.0f;
            } else {
                uint32_t r2[32], r0[32];
                tmem_ld32(lane_base + TM_DP + 32, r2);              // tap 2: g2[j] . W_k2
                tmem_ld32(lane_base + TM_DP + 64, r0);              // tap 0: g2[j] . W_k0, needed by position 2(j-1)+1
                tmem_ld_wait();
                float *xc = xchg + lq * 32;
                if (lane == 0) {
#pragma unroll
                    for (int c1 = 0; c1 < 32; ++c1) xc[c1] = __uint_as_float(r0[c1]);
                }
                named_bar_sync(2, NWORK / 2);
#pragma unroll
                for (int c1 = 0; c1 < 32; ++c1) {
                    float dn = __uint_as_float(__shfl_down_sync(0xffffffffu, r0[c1], 1));
                    if (lane == 31) dn = lq < 3 ? xc[32 + c1] : 0.0f;       // row j+1 lives in the next warp (j = 127: none)
                    g1[c1] = ((hmask >> c1) & 1u) ? __uint_as_float(r2[c1]) + dn : 0.0f;
                }
            }
            // ---- g1T and im2colT overwrite HT / A1: P2 (and conv1) must have retired
            mbar_wait(&bars[B_P2], ph);
#pragma unroll
            for (int c1 = 0; c1 < 32; ++c1) {
                const int idx = ka * G1_ATOM + sw128_index(half * 32 + c1, jc);
                const float hi = tf32_hi(g1[c1]);
                ht[idx] = hi;
                ht[G1_LO + idx] = g1[c1] - hi;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int idx = ka * IM_ATOM + sw128_index(half * 16 + k, jc);
                const float hi = tf32_hi(v[k]);
                a1[idx] = hi;
                a1[IM_LO + idx] = v[k] - hi;
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars[B_G1]);
        }
        // =================================================================== per-CTA partial -> global
        mbar_wait(&bars[B_P3], (uint32_t)(it - 1) & 1u);
        tc_fence_after();
        float *stage = gt;                                         // PART floats (G / g2T are dead)
        float *red = gq;                                           // [256][16] conv2-bias partials
        uint32_t r1[32];
        if (half == 0) {
            uint32_t r2[32];
            tmem_ld32(lane_base + TM_ACC2, r2);
            tmem_ld32(lane_base + TM_ACC1, r1);
            tmem_ld_wait();
            if (j < 96) {
                const int blk = j >> 5, c1 = j & 31, tap = blk == 0 ? 1 : (blk == 1 ? 2 : 0);
#pragma unroll
                for (int co = 0; co < 32; ++co) stage[co * 96 + c1 * 3 + tap] = __uint_as_float(r2[co]);
            }
            if (j < 32) {                                          // even half of dW1 | db1
#pragma unroll
                for (int k = 0; k < 15; ++k) stage[3104 + j * 15 + k] = __uint_as_float(r1[k]);
                stage[3584 + j] = __uint_as_float(r1[15]);
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) red[tid * 16 + c] = acc_b2[c];
        named_bar_sync(1, NWORK);
        if (half == 0 && j >= 32 && j < 64) {                      // odd half: rows 32..63, columns 16..31
            const int c1 = j - 32;
#pragma unroll
            for (int k = 0; k < 15; ++k) stage[3104 + c1 * 15 + k] += __uint_as_float(r1[16 + k]);
            stage[3584 + c1] += __uint_as_float(r1[31]);
        }
        if (tid < 32) {                                            // db2[co] = sum over the 128 positions
            const int hsel = tid >> 4, c = tid & 15;
            float sacc = 0.0f;
            for (int q = 0; q < 128; ++q) sacc += red[(hsel * 128 + q) * 16 + c];
            stage[3072 + tid] = sacc;
        }
        named_bar_sync(1, NWORK);
        float *dst = part + ((size_t)t * slots + slot) * PART;
        for (int i = tid; i < PART; i += NWORK) dst[i] = stage[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, bw::TM_COLS);
    }
}

}  // namespace

int rlca_conv_tc_init()
{
    cudaError_t e = cudaFuncSetAttribute(conv_tower_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess)
        return rlca_set_err(RLCA_ERR_CUDA, "cudaFuncSetAttribute(conv_tower_fwd_tc_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(conv_tower_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bw::SMEM_BYTES);
    if (e != cudaSuccess)
        return rlca_set_err(RLCA_ERR_CUDA, "cudaFuncSetAttribute(conv_tower_bwd_tc_kernel): %s", cudaGetErrorString(e));
    return RLCA_OK;
}

size_t rlca_conv_tc_image_floats() { return (size_t)W_FLOATS; }

void rlca_conv_tc_prep(const float *const cv1w[2], const float *const cv1b[2], const float *const cv2w[2],
                       const float *const cv2b[2], float *img, cudaStream_t s)
{
    ConvTcWeights w;
    for (int t = 0; t < 2; ++t) { w.cv1w[t] = cv1w[t]; w.cv1b[t] = cv1b[t]; w.cv2w[t] = cv2w[t]; w.cv2b[t] = cv2b[t]; }
    conv_tc_prep_kernel<<<(W_FLOATS + 255) / 256, 256, 0, s>>>(w, img);
}

int rlca_conv_tc_forward(const float *obs, const float *img, float *F, float *Fs, int nb, int num_sms, cudaStream_t s)
{
    const int grid = nb < num_sms ? nb : num_sms;
    conv_tower_fwd_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, s>>>(obs, img, F, Fs, nb);
    RLCA_CUDA_TRY(cudaGetLastError());
    return RLCA_OK;
}

size_t rlca_conv_tc_bwd_image_floats() { return 2 * (size_t)bw::WB_TOWER; }

int rlca_conv_tc_bwd_slots(int nb, int num_sms)
{
    const int per_tower = num_sms / 2 > 0 ? num_sms / 2 : 1;
    return nb < per_tower ? nb : per_tower;
}

void rlca_conv_tc_bwd_prep(const float *const cv1w[2], const float *const cv1b[2], const float *const cv2w[2],
                           const float *const cv2b[2], float *img, cudaStream_t s)
{
    ConvTcWeights w;
    for (int t = 0; t < 2; ++t) { w.cv1w[t] = cv1w[t]; w.cv1b[t] = cv1b[t]; w.cv2w[t] = cv2w[t]; w.cv2b[t] = cv2b[t]; }
    conv_tc_bwd_prep_kernel<<<dim3((bw::WB_TOWER + 255) / 256, 2), 256, 0, s>>>(w, img);
}

int rlca_conv_tc_backward(const float *obs, const float *img, const float *dF, const float *Fmask, float *part, int nb,
                          int num_sms, cudaStream_t s)
{
    const int slots = rlca_conv_tc_bwd_slots(nb, num_sms);
    conv_tower_bwd_tc_kernel<<<2 * slots, NTHREADS, bw::SMEM_BYTES, s>>>(obs, img, dF, Fmask, part, nb, slots);
    RLCA_CUDA_TRY(cudaGetLastError());
    return RLCA_OK;
}
