// Probe: tcgen05.mma.kind::tf32 with MN-major (transposed) shared-memory operands, 128B swizzle.
// D[128 x 32] = sum_k A[k][m] * B[k][n], A stored as 4 tiles [K rows x 32 m] (tile stride = LBO), B as one tile [K rows x 32 n].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I rl_collision_avoidance_b200/csrc -o tools/probes/probe_mn_major tools/probes/probe_mn_major.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "rlca_tc_ptx.cuh"
using namespace rlca_ptx;

constexpr int KTOT = 32;                 // 4 MMAs of K = 8
constexpr int TILE = 128 * 32;           // floats per [128 rows x 32] tile (rows >= KTOT used)

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void probe(const float *A, const float *B, float *D, int variant)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = raw + ((1024u - (smem_addr(raw) & 1023u)) & 1023u);
    float *sa = reinterpret_cast<float *>(sm);            // 4 tiles
    float *sb = sa + 4 * TILE;                            // 1 tile
    uint64_t *bar = reinterpret_cast<uint64_t *>(sb + 3 * TILE);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // A[k][m] -> tile m/32, row k, col m%32 (swizzled)
    for (int i = tid; i < KTOT * 128; i += blockDim.x) {
        const int k = i / 128, m = i % 128;
        sa[(m >> 5) * TILE + sw128_index(k, m & 31)] = A[i];
    }
    for (int i = tid; i < KTOT * 32; i += blockDim.x) {
        const int k = i / 32, n = i % 32;
        sb[sw128_index(k, n)] = B[i];
    }
    // K-major copies (variants 2..4): A -> [128 m rows x 32 k], B -> [32 n rows x 32 k]
    float *sak = sb + TILE, *sbk = sak + TILE;
    for (int i = tid; i < KTOT * 128; i += blockDim.x) { const int k = i / 128, m = i % 128; sak[sw128_index(m, k)] = A[i]; }
    for (int i = tid; i < KTOT * 32; i += blockDim.x) { const int k = i / 32, n = i % 32; sbk[sw128_index(n, k)] = B[i]; }
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tptr, 32);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    if (tid == 0) {
        // idesc: tf32 x tf32 -> f32, a_major (bit 15) = b_major (bit 16) = 1 (MN-major), N = 32, M = 128
        const bool a_mn = variant == 0 || variant == 1 || variant == 4, b_mn = variant == 0 || variant == 1 || variant == 3;
        const uint32_t idesc = umma_idesc_tf32(128, 32) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u);
        for (int kb = 0; kb < KTOT / 8; ++kb) {
            const uint32_t lbo = variant == 1 ? 1024 : TILE * 4, sbo = variant == 1 ? TILE * 4 : 1024;
            const uint64_t ad = a_mn ? desc_mn(smem_addr(sa) + kb * 1024, lbo, sbo) : umma_desc_sw128(smem_addr(sak) + kb * 32);
            const uint64_t bd = b_mn ? desc_mn(smem_addr(sb) + kb * 1024, lbo, sbo) : umma_desc_sw128(smem_addr(sbk) + kb * 32);
            umma_tf32(tmem, ad, bd, idesc, kb ? 1u : 0u);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (warp < 4) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
        for (int n = 0; n < 32; ++n) D[(warp * 32 + lane) * 32 + n] = __uint_as_float(r[n]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

int main()
{
    std::vector<float> A(KTOT * 128), B(KTOT * 32), D(128 * 32), ref(128 * 32, 0.f);
    srand(1);
    auto rnd = [] { float x = (rand() % 2001 - 1000) / 1000.0f; uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
    for (auto &x : A) x = rnd();
    for (auto &x : B) x = rnd();
    for (int k = 0; k < KTOT; ++k)
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) ref[m * 32 + n] += A[k * 128 + m] * B[k * 32 + n];
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 7 * TILE * 4 + 1024 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int variant = 0; variant < 5; ++variant) {
        cudaMemset(dD, 0, D.size() * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0; int bad_row = -1;
        for (int i = 0; i < 128 * 32; ++i) { double d = fabs(D[i] - ref[i]); if (d > err) { err = d; bad_row = i / 32; } }
        // per 32-row block error to see which atoms are right
        printf("variant %d (%s): %s max err %.3e (row %d);", variant, variant == 0 ? "A,B MN-major: LBO=MN tile stride, SBO=K group" : variant == 1 ? "A,B MN-major swapped" : variant == 2 ? "A,B K-major" : variant == 3 ? "A K-major, B MN-major" : "A MN-major, B K-major",
               cudaGetErrorString(e), err, bad_row);
        for (int blk = 0; blk < 4; ++blk) {
            double eb = 0;
            for (int i = blk * 1024; i < (blk + 1) * 1024; ++i) eb = fmax(eb, fabs(D[i] - ref[i]));
            printf(" blk%d %.2e", blk, eb);
        }
        printf(" D[0..3]= %.3f %.3f %.3f %.3f ref %.3f %.3f\n", D[0], D[1], D[2], D[3], ref[0], ref[1]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
