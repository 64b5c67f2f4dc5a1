// Experiment: does a K-major SWIZZLE_128B A-operand descriptor work with a start address that is 128-byte but not 1024-byte
// aligned (rows shifted by s), when the tile was written with the absolute-address swizzle?  With and without base_offset.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I diffusion_conductor_b200/csrc -o gpurun_out/desc_shift tools/experiments/desc_shift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace dc;

__global__ void k(const uint16_t* a /*[272][64]*/, const uint16_t* b /*[16][64]*/, float* out /*[2][9][128][16]*/) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    uint8_t* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
    uint8_t* A = sm;                 // 272 rows x 128 B, chunk c of row r at c ^ (r & 7)  (r = absolute row, base 1024-aligned)
    uint8_t* B = sm + 272 * 128 + 1024 - ((272 * 128) & 1023);
    B = sm + 36864;                  // 1024-aligned, beyond A (34816)
    __shared__ uint64_t bar;
    __shared__ uint32_t tb;
    const int tid = threadIdx.x;
    for (int i = tid; i < 272 * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(A + sw128_offset(r, c)) = reinterpret_cast<const uint4*>(a)[i];
    }
    for (int i = tid; i < 16 * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(B + sw128_offset(r, c)) = reinterpret_cast<const uint4*>(b)[i];
    }
    if (tid == 0) mbar_init(smem_u32(&bar), 1), mbar_fence_init();
    if (tid < 32) tmem_alloc(smem_u32(&tb), 32), tmem_relinquish();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tb;
    uint32_t ph = 0;
    for (int mode = 0; mode < 2; ++mode)
        for (int s = 0; s < 9; ++s) {
            if (tid == 0) {
                const uint32_t start = smem_u32(A) + s * 128;
                uint64_t ad = make_desc_kmajor_sw128(start);
                if (mode == 1) ad |= (uint64_t)((start >> 7) & 7u) << 49;
                const uint64_t bd = make_desc_kmajor_sw128(smem_u32(B));
                const uint32_t idesc = make_idesc<true>(128, 16);
                for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, ad + 2 * kk, bd + 2 * kk, idesc, kk > 0);
                umma_commit(smem_u32(&bar));
            }
            mbar_wait(smem_u32(&bar), ph & 1u);
            ++ph;
            tc_fence_after();
            if (tid < 128) {
                float v[16];
                tmem_ld16(tmem + ((uint32_t)((tid >> 5) * 32) << 16), v);
                tmem_wait_ld();
                for (int j = 0; j < 16; ++j) out[(((size_t)mode * 9 + s) * 128 + tid) * 16 + j] = v[j];
            }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        }
    if (tid < 32) tmem_dealloc(tmem, 32);
}

static uint16_t bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)((u + 0x7FFF + ((u >> 16) & 1)) >> 16); }
static float fb(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
int main() {
    std::vector<uint16_t> a(272 * 64), b(16 * 64);
    srand(1);
    for (auto& v : a) v = bf((rand() % 200 - 100) / 64.f);
    for (auto& v : b) v = bf((rand() % 200 - 100) / 64.f);
    uint16_t *da, *db; float* dout;
    cudaMalloc(&da, a.size() * 2), cudaMalloc(&db, b.size() * 2), cudaMalloc(&dout, 2 * 9 * 128 * 16 * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice), cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<1, 128, 65536>>>(da, db, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> out(2 * 9 * 128 * 16);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    for (int mode = 0; mode < 2; ++mode)
        for (int s = 0; s < 9; ++s) {
            double maxerr = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < 16; ++n) {
                    double ref = 0;
                    for (int kk = 0; kk < 64; ++kk) ref += (double)fb(a[(r + s) * 64 + kk]) * fb(b[n * 64 + kk]);
                    maxerr = fmax(maxerr, fabs(ref - out[(((size_t)mode * 9 + s) * 128 + r) * 16 + n]));
                }
            printf("base_offset %s, row shift %d: max err %.3g\n", mode ? "set" : "0  ", s, maxerr);
        }
    return 0;
}
