// umma_pipe.cu -- variants of umma_rate.cu that add, one at a time, what the real kernels do around the MMA chain:
// a tcgen05.commit after every k-block of 12 MMAs, and other warps streaming data into shared memory meanwhile.
// umma_rate.cu -- how long does ONE tcgen05.mma.kind::tf32 (K = 8) take as a function of its M and N?
// One CTA per SM issues a long chain of MMAs on the same shared-memory operands (SWIZZLE_128B, K-major) into the same
// tensor-memory accumulator, commits, waits, and reports SM cycles per instruction.  The recurrent sweep (sweep_tc.cu)
// is a chain of small-N MMAs; this measures the per-instruction floor that decides which operand should be the
// recurrent matrix.          build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// nops: MMAs per k-block pattern: distinct A/B sub-tiles are cycled the way the sweep does (4 k-steps x 3 products)
__global__ void __launch_bounds__(256) rate_kernel(int M, int N, int iters, int commit_every, int writers, long long *out) {
    const int same_operands = 0;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *base = (uint8_t *)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t kbar[4];
    __shared__ volatile int stop;
    __shared__ uint32_t slot;
    float *f = (float *)base;
    for (int i = threadIdx.x; i < (2 * 128 * 32 + 2 * 256 * 32); i += blockDim.x) f[i] = 0.001f * (float)(i % 97);
    if (threadIdx.x == 0) {
        stop = 0;
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&kbar[i])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(base), a1 = a0 + 128 * 128, b0 = a0 + 2 * 128 * 128, b1 = b0 + 256 * 128;
        const uint64_t dA = smem_desc(a0), dAl = smem_desc(a1), dB = smem_desc(b0), dBl = smem_desc(b1);
        const uint32_t id = idesc(M, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
                const uint64_t adv = same_operands ? 0ull : (uint64_t)((k8 * 32) >> 4);
                const uint64_t xa[3] = {dAl + adv, dA + adv, dA + adv}, xb[3] = {dB + adv, dBl + adv, dB + adv};
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem), "l"(xa[p]), "l"(xb[p]), "r"(id), "r"(1u) : "memory");
                }
            }
            if (commit_every) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&kbar[it & 3])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
        stop = 1;
    } else if (writers && threadIdx.x >= 128) {
        // what TMA + splitter traffic looks like to the shared-memory port: warps 4-7 stream 16-byte stores into a
        // scratch region (96 KB behind the operands) until the MMA thread is done
        float4 *scr = reinterpret_cast<float4 *>(base + 98304);
        int i = threadIdx.x - 128;
        while (!stop) {
            scr[i & 4095] = make_float4(1.f, 2.f, 3.f, 4.f);
            i += 128;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

int main() {
    long long *d;
    cudaMalloc(&d, 8);
    const size_t smem = (2 * 128 * 128 + 2 * 256 * 128) + 65536 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("tcgen05.mma.kind::tf32 K=8 M=128, chain of 12 x 400 instructions per SM, all %d SMs busy\n", sms);
    const int Ns[] = {32, 64, 256};
    for (int writers = 0; writers < 2; ++writers)
        for (int ce = 0; ce < 2; ++ce)
            for (int N : Ns) {
                const int iters = 400;
                rate_kernel<<<sms, 256, smem>>>(128, N, 4, ce, writers, d);
                rate_kernel<<<sms, 256, smem>>>(128, N, iters, ce, writers, d);
                long long h = 0;
                cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
                printf("  N=%3d  commit per 12 MMAs: %d  4 warps storing to smem: %d  : %7.1f cycles/MMA\n", N, ce, writers, (double)h / (12.0 * iters));
            }
    return 0;
}
