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

}  // namespace

int rlca_conv_tc_init()
{
    cudaError_t e = cudaFuncSetAttribute(conv_tower_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess)
        return rlca_set_err(RLCA_ERR_CUDA, "cudaFuncSetAttribute(conv_tower_fwd_tc_kernel): %s", cudaGetErrorString(e));
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
