// umma_layout_probe.cu -- which shared-memory word does tcgen05.mma read for A(m, k) under a given descriptor?
// A's 16 KB region is filled with its own word index; B (K-major, N = 16) is the 8 x 8 identity, so D[m][n] = A(m, k = n)
// = the word index the tensor core fetched.  Used to pin down the MN-major / SWIZZLE_128B descriptor of gemm_tc.cu.
// tf32 keeps 11 significant bits, so the index is sent in two passes (low 11 bits, then the rest).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_layout_probe umma_layout_probe.cu
//   run:   ./umma_layout_probe <a_major 0|1> <LBO bytes> <SBO bytes> <layout_type 0|2|4|6> [start offset bytes]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(int a_major, int lbo, int sbo, int layout, int start_off, int pass, float *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *base = (uint8_t *)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    float *A = (float *)base;                      // 16 KB probe region
    float *B = (float *)(base + 32768);            // N = 16 rows x 128 B (K-major, SWIZZLE_128B): identity in k = 0..7
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) A[i] = (pass == 0) ? (float)(i & 2047) : (float)(i >> 11);
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) B[i] = 0.f;
    __syncthreads();
    if (threadIdx.x < 8) {       // B[n][k] = 1 for k == n: K-major row n, 16-byte chunk (k/4) XOR (n % 8) under the 128B swizzle
        const int n = threadIdx.x, k = threadIdx.x;
        const int chunk = (k / 4) ^ (n % 8);
        B[n * 32 + chunk * 4 + (k % 4)] = 1.f;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        uint64_t dA = 0, dB = 0;
        const uint32_t aa = smem_u32(base) + (uint32_t)start_off, ab = smem_u32(B);
        dA |= (uint64_t)((aa & 0x3ffff) >> 4);
        dA |= (uint64_t)((uint32_t)lbo >> 4) << 16;
        dA |= (uint64_t)((uint32_t)sbo >> 4) << 32;
        dA |= (uint64_t)1 << 46;
        dA |= (uint64_t)layout << 61;
        dB |= (uint64_t)((ab & 0x3ffff) >> 4);
        dB |= (uint64_t)1 << 16;
        dB |= (uint64_t)(1024 >> 4) << 32;
        dB |= (uint64_t)1 << 46;
        dB |= (uint64_t)2 << 61;
        const uint32_t id = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_major & 1) << 15) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(dA), "l"(dB), "r"(id), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        uint32_t r[8];
        const int warp = threadIdx.x >> 5;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[threadIdx.x * 8 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
    }
}

int main(int argc, char **argv) {
    const int a_major = argc > 1 ? atoi(argv[1]) : 1, lbo = argc > 2 ? atoi(argv[2]) : 4096, sbo = argc > 3 ? atoi(argv[3]) : 1024;
    const int layout = argc > 4 ? atoi(argv[4]) : 2, start = argc > 5 ? atoi(argv[5]) : 0;
    float *d;
    cudaMalloc(&d, 2 * 128 * 8 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    for (int pass = 0; pass < 2; ++pass) probe<<<1, 128, 48 * 1024>>>(a_major, lbo, sbo, layout, start, pass, d + pass * 1024);
    float h[2048];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    printf("a_major=%d LBO=%d SBO=%d layout=%d start=%d : word index (byte offset = 4x) read for A(m, k)\n", a_major, lbo, sbo, layout, start);
    const int rows[] = {0, 1, 2, 7, 8, 9, 31, 32, 33, 63, 64, 96, 127};
    for (int m : rows) {
        printf("  m=%3d :", m);
        for (int k = 0; k < 8; ++k) printf(" %5d", (int)h[m * 8 + k] + 2048 * (int)h[1024 + m * 8 + k]);
        printf("\n");
    }
    return 0;
}
