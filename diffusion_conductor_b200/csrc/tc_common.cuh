// sm_100a building blocks: mbarrier, bulk async copy (TMA unit, non-tensor form), TMEM
// allocation / load / store, tcgen05.mma with shared-memory matrix descriptors.
//
// Layout convention used everywhere in this library ("K-major SW128 image"):
//   an operand block is [rows x 64] 16-bit elements, one row = 128 bytes, rows contiguous;
//   inside each 8-row group (1024 B) the 16-byte chunk c of row r lives at chunk c ^ (r & 7).
//   This is exactly what tcgen05.mma expects for a K-major SWIZZLE_128B operand with
//   SBO = 1024 B, so a block that was written to global memory in this image can be brought
//   in with one linear cp.async.bulk -- no tensor map needed.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dc {

constexpr int kTileRows = 128;            // tokens per tile == TMEM lanes
constexpr int kBlockK = 64;               // 16-bit elements per 128-byte swizzle row
constexpr int kABlockBytes = kTileRows * 128;   // one [128 x 64] operand block

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- programmatic dependent launch
// pdl_trigger(): let the next kernel in the stream be scheduled as soon as every CTA of this one has started;
// pdl_wait(): block until the previous kernel has completed and its writes are visible.  Both are no-ops when
// the kernel was launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// non-blocking probe
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// Blocking wait: the whole retry loop is two instructions (try_wait suspends the thread in hardware up to the
// time hint and wakes it when the phase completes), so waiting warps leave the issue slots to the working ones.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
}

// ---------------------------------------------------------------- bulk async copy global -> smem
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// same copy issued as independent pieces of kBulkChunk bytes (experiment knob: the TMA unit serves several
// outstanding bulk requests concurrently, one big request is served sequentially)
#ifndef DC_BULK_CHUNK
#define DC_BULK_CHUNK 4096
#endif
__device__ __forceinline__ void bulk_g2s_chunked(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    const uint8_t* p = static_cast<const uint8_t*>(src);
    for (uint32_t off = 0; off < bytes; off += DC_BULK_CHUNK)
        bulk_g2s(dst_smem + off, p + off, min((uint32_t)DC_BULK_CHUNK, bytes - off), bar);
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 4 consecutive columns
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 8 consecutive columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

// 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}

// named barrier among a subset of warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- tcgen05.mma
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
//   bits [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, canonical value 1)
//   bits [32,46) SBO>>4 | [46,48) version=1 (sm_100) | [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// MN-major (transposed) operand, SWIZZLE_128B: the image is the same [rows x 64]-block image, read with the
// block's 128-byte rows as the K index (8 rows = one K atom, 1024 B apart -> SBO) and the 64 elements of a row
// as the M/N index (blocks of 64 are kABlockBytes apart -> LBO).  Used for K^T V, where the contraction runs
// over tokens.  One MMA consumes K = 16 = two K atoms; advance the start address by 2048 B per step.
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(kABlockBytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
template <bool kBf16>
__device__ __forceinline__ constexpr uint32_t make_idesc_mn(int M, int N) {   // both operands MN-major
    return (1u << 4) | ((kBf16 ? 1u : 0u) << 7) | ((kBf16 ? 1u : 0u) << 10) | (1u << 15) | (1u << 16) |
           (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// order-preserving float <-> int key (for redux.sync max on floats)
__device__ __forceinline__ int f2key(float x) {
    const int i = __float_as_int(x);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

// Instruction descriptor for kind::f16, fp32 accumulate, both operands K-major.
//   [4,6) D fmt (1=f32) | [7,10) A fmt | [10,13) B fmt (0=f16, 1=bf16) | [17,23) N>>3 | [24,29) M>>4
template <bool kBf16>
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((kBf16 ? 1u : 0u) << 7) | ((kBf16 ? 1u : 0u) << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one elected thread issues.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, with a 128-bit "disable output lane" mask: bit r set -> row r (TMEM lane r) of D is left
// untouched.  Lets one A operand be multiplied by a different B per row segment.
__device__ __forceinline__ void umma_f16_masked(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                                const uint32_t* mask) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(mask[0]), "r"(mask[1]), "r"(mask[2]), "r"(mask[3])
        : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// One [128 x 64] x [N x 64]^T k-block = 4 instructions of K=16 (32 bytes along the swizzled row).
__device__ __forceinline__ void umma_kblock(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem, uint32_t idesc, bool accumulate_first) {
    const uint64_t ad = make_desc_kmajor_sw128(a_smem);
    const uint64_t bd = make_desc_kmajor_sw128(b_smem);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (accumulate_first || k > 0) ? 1u : 0u);
}

__device__ __forceinline__ void umma_kblock_masked(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem, uint32_t idesc, bool accumulate_first,
                                                   const uint32_t* mask) {
    const uint64_t ad = make_desc_kmajor_sw128(a_smem);
    const uint64_t bd = make_desc_kmajor_sw128(b_smem);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16_masked(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (accumulate_first || k > 0) ? 1u : 0u, mask);
}

// ---------------------------------------------------------------- 2-CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// ---------------------------------------------------------------- Ampere-style async copy global -> smem (no register staging)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------- distributed shared memory (one cluster = one clip)
// Writer: plain st.shared of the payload, CTA barrier, then one thread per peer arrives on the peer's mbarrier with
// release.cluster (cumulative over the barrier: no separate fence.acq_rel.cluster, which costs an L1 flush).  Reader: one
// thread waits with acquire.cluster, CTA barrier, then everyone pulls with ld.shared::cluster.
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_acq_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}"
        ::"r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void st_dsmem_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_dsmem_f32x4(uint32_t cluster_addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ float2 ld_dsmem_f32x2(uint32_t cluster_addr) {
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_addr) : "memory");
    return v;
}

// ---------------------------------------------------------------- flag-in-data exchange through global memory (L2)
// Every 8-byte word is (value, tag): a naturally aligned 8-byte store is performed as one access, so a reader that sees the
// expected tag also sees the value -- no fence, no separate flag.  Volatile accesses bypass L1.
__device__ __forceinline__ void st_global_v2(uint2* p, uint2 v) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_global_v4(uint4* p, uint4 v) {          // two (value, tag) words
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// The barriers the MMA issuer waits on are local to the leader CTA; the peer arrives on them with
// release.cluster after fencing its own shared-memory writes for the async proxy.  Those writes are read by the
// peer SM's own tensor-core datapath, so the ordinary CTA-scope probe is what is needed here (cluster-scope
// acquires / fences measured 1.5-2x slower for the whole kernel: they invalidate L1).
__device__ __forceinline__ bool mbar_test_cluster(uint32_t bar, uint32_t parity) { return mbar_test(bar, parity); }
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's 128 rows] * B[N/2 rows from each CTA]^T, issued by the leader CTA.
// mask: 256-bit "disable output lane" (rows 0..127 = leader CTA, 128..255 = peer).
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                              const uint32_t* mask) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(mask[0]), "r"(mask[1]), "r"(mask[2]), "r"(mask[3]),
        "r"(mask[4]), "r"(mask[5]), "r"(mask[6]), "r"(mask[7])
        : "memory");
}
__device__ __forceinline__ void umma_kblock_2cta(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem, uint32_t idesc, bool accumulate_first,
                                                 const uint32_t* mask) {
    const uint64_t ad = make_desc_kmajor_sw128(a_smem);
    const uint64_t bd = make_desc_kmajor_sw128(b_smem);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16_2cta(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (accumulate_first || k > 0) ? 1u : 0u, mask);
}
// arrives on the mbarrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

// ---------------------------------------------------------------- packed fp32x2 math (FFMA2 / FADD2 / FMUL2)
// Two fp32 lanes per instruction: halves the issue slots of the row-wise math, which is issue-bound.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---------------------------------------------------------------- operand element type
template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    if constexpr (kBf16) {
        __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&p);
    } else {
        __half2 p = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&p);
    }
}
template <bool kBf16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
    if constexpr (kBf16) {
        return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
    } else {
        return __half22float2(*reinterpret_cast<__half2*>(&u));
    }
}
template <bool kBf16>
__device__ __forceinline__ uint16_t pack1(float v) {
    if constexpr (kBf16) {
        __nv_bfloat16 p = __float2bfloat16_rn(v);
        return *reinterpret_cast<uint16_t*>(&p);
    } else {
        __half p = __float2half_rn(v);
        return *reinterpret_cast<uint16_t*>(&p);
    }
}

// byte offset of (row r, 16-byte chunk c) inside a [rows x 64] K-major SW128 block
__device__ __host__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

// Compact block-diagonal attention image (persistent kernel): only the eight 16 x 16 head blocks of A = softmax_T(K)^T V exist.
// Two [16 rows x 128 B] SW128 blocks; block hh / 4, row l (value column of the head), K elements 16 (hh % 4) + d (key feature of the
// head): y[:, 16 hh + l] = sum_d q[:, 16 hh + d] A_hh[d][l] is ONE M = 128, N = 16, K = 16 MMA per head (an M = 128 MMA costs the same
// ~64 cycles for N = 16 as for N = 128), and the image is 4 KB instead of a 32 KB block-diagonal [128 x 128] one that is 7/8 zeros.
constexpr uint32_t kBdcBytes = 4096;
__device__ __host__ __forceinline__ uint32_t bdc_offset(uint32_t hh, uint32_t d, uint32_t l) {
    const uint32_t k = (hh & 3u) * 16u + d;
    return (hh >> 2) * 2048u + l * 128u + (((k >> 3) ^ (l & 7u)) << 4) + (k & 7u) * 2u;
}

// write 16 consecutive fp32 (columns col0..col0+15 of the 128-wide row, col0 % 16 == 0) of row r
// into the A-operand buffer made of [128 x 64] blocks.
template <bool kBf16>
__device__ __forceinline__ void store_a16(uint32_t a_base, uint32_t r, uint32_t col0, const float* v) {
    const uint32_t kb = col0 >> 6, c = (col0 & 63u) >> 3;
    uint32_t p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pack2<kBf16>(v[2 * i], v[2 * i + 1]);
    const uint32_t a0 = a_base + kb * kABlockBytes + sw128_offset(r, c);
    const uint32_t a1 = a_base + kb * kABlockBytes + sw128_offset(r, c + 1);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]) : "memory");
}

}  // namespace dc
