// umma_rate - how fast does one SM retire tcgen05.mma when the operands stream from shared memory?
//
// The denoiser's level-0 convs are M = 128, N = 32 MMAs whose A operand (4 KB) and B operand (1 KB) are read from shared memory through
// NO-SWIZZLE K-major descriptors (dn_conv_tc.cuh: the halo-tile trick needs 16-byte tap offsets).  ncu on enc1.l2a: ~80-100 cycles per
// MMA against a 16-cycle tensor floor, i.e. ~64 B/clk of operand bandwidth.  Before redesigning the activation layout this probe answers,
// per SM and with nothing else running:
//   1. cycles per MMA for N = 16 .. 256 with the conv engine's descriptors (SBO 160 B, LBO 2880 B),
//   2. the same with dense no-swizzle core matrices and with SWIZZLE_128B operands (64 fp16 channels per 128-byte row),
//   3. the same with the A operand in TMEM (tcgen05.mma [d], [a], b-desc): the floor when only B streams from shared memory.
// Numerics are irrelevant (shared memory is zero filled); only descriptor validity and timing matter.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/umma_rate tools/microbench/umma_rate.cu
//   run:   tools/microbench/umma_rate [iters=2048]      (one line per configuration; run it under `timeout`)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct Cfg {
    int n;                    // MMA N (multiple of 8 for M = 128 at cta_group::1... 16 is the safe granularity)
    int layout;               // UMMA layout type: 0 = no swizzle, 2 = SWIZZLE_128B
    uint32_t a_lbo, a_sbo;    // bytes
    uint32_t b_lbo, b_sbo;
    uint32_t a_kstep, b_kstep;// bytes the start address advances per K step inside one staged tile
    int ksteps;               // K steps per staged tile
    uint32_t a_tile, b_tile;  // bytes between the tiles the loop rotates through (streaming: a new tile every `per_tile` MMAs)
    int a_tiles, b_tiles;     // how many tiles fit in the A region (96 KB) / B region (64 KB)
    int per_tile;             // MMAs issued per staged tile (conv engine: 18 = 9 taps x 2 K steps)
    int a_in_tmem;            // A operand from TMEM instead of shared memory
    int tf32;                 // kind::tf32 (K = 8) instead of kind::f16 (K = 16); operand bytes per MMA are the same
    int iters;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) umma_rate_kernel(const Cfg c, unsigned long long* cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy zero fill -> async-proxy (tensor core) reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t fmt = c.tf32 ? 2u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(c.n >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        unsigned long long t0 = 0, t1 = 0;
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
        if (pred) {
            t0 = clock64();
            int in_tile = 0, ks = 0, at = 0, bt = 0;                      // no divisions in the issue loop: it must outrun a 16-cycle MMA
            for (int i = 0; i < c.iters; ++i) {
                const uint32_t a_addr = a0 + (uint32_t)at * c.a_tile + (uint32_t)ks * c.a_kstep;
                const uint32_t b_addr = b0 + (uint32_t)bt * c.b_tile + (uint32_t)ks * c.b_kstep;
                const uint64_t bd = make_desc(b_addr, c.b_lbo, c.b_sbo, c.layout);
                if (c.a_in_tmem) {
                    const uint32_t a_t = tmem + 256;                          // M = 128 lanes x 8 columns (16 fp16 / 8 tf32 per lane)
                    if (c.tf32) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(a_t), "l"(bd), "r"(idesc), "r"(i ? 1 : 0) : "memory");
                    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(a_t), "l"(bd), "r"(idesc), "r"(i ? 1 : 0) : "memory");
                } else {
                    const uint64_t ad = make_desc(a_addr, c.a_lbo, c.a_sbo, c.layout);
                    if (c.tf32) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1 : 0) : "memory");
                    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1 : 0) : "memory");
                }
                if (++ks == c.ksteps) ks = 0;
                if (++in_tile == c.per_tile) { in_tile = 0; if (++at == c.a_tiles) at = 0; if (++bt == c.b_tiles) bt = 0; }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        // every lane waits for the MMAs to retire (phase 0)
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        if (pred) { t1 = clock64(); cycles_out[blockIdx.x] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
    (void)lane;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2048;
    int dev = 0, sms = 0;
    CHECK(cudaSetDevice(dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CHECK(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024));
    unsigned long long* d_cycles = nullptr;
    CHECK(cudaMalloc(&d_cycles, sizeof(unsigned long long) * sms));
    struct Named { const char* name; Cfg c; };
    const int Ns[] = {16, 32, 48, 64, 112, 128, 256};
    printf("# %d SMs, %d MMAs per CTA, M = 128, 32 operand bytes per row per MMA (K = 16 fp16 / 8 tf32); grid = 1 CTA and 1 CTA per SM\n", sms, iters);
    printf("%-34s %5s %6s %12s %12s %10s %10s\n", "operands", "N", "kind", "cyc/MMA(1)", "cyc/MMA(all)", "B/clk(1)", "floor");
    for (int variant = 0; variant < 5; ++variant)
        for (int tf32 = 0; tf32 < 2; ++tf32)
            for (int n : Ns) {
                Cfg c = {};
                c.n = n; c.iters = iters; c.tf32 = tf32; c.per_tile = 18;
                const char* name = "";
                const uint32_t nb = (uint32_t)n;
                if (variant == 0) { name = "conv engine: no-swizzle halo tile"; c.layout = 0; c.a_lbo = 2880; c.a_sbo = 160; c.a_kstep = 2 * 2880; c.a_tile = 11520;
                                    c.b_lbo = nb * 16; c.b_sbo = 128; c.b_kstep = nb * 32; c.b_tile = nb * 64; c.ksteps = 2; c.per_tile = 18; }
                if (variant == 1) { name = "dense no-swizzle core matrices"; c.layout = 0; c.a_lbo = 128; c.a_sbo = 256; c.a_kstep = 0; c.a_tile = 4096;
                                    c.b_lbo = 128; c.b_sbo = 256; c.b_kstep = 0; c.b_tile = nb * 32; c.ksteps = 1; c.per_tile = 1; }
                if (variant == 2) { name = "SWIZZLE_128B rows of 64 fp16"; c.layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_kstep = 32; c.a_tile = 16384;
                                    c.b_lbo = 16; c.b_sbo = 1024; c.b_kstep = 32; c.b_tile = nb * 128; c.ksteps = 4; c.per_tile = 4; }
                if (variant == 3) { name = "A in TMEM, B no-swizzle (conv)"; c.layout = 0; c.a_in_tmem = 1; c.a_tile = 4096;
                                    c.b_lbo = nb * 16; c.b_sbo = 128; c.b_kstep = nb * 32; c.b_tile = nb * 64; c.ksteps = 2; c.per_tile = 18; }
                if (variant == 4) { name = "A in TMEM, B SWIZZLE_128B"; c.layout = 2; c.a_in_tmem = 1; c.a_tile = 4096;
                                    c.b_lbo = 16; c.b_sbo = 1024; c.b_kstep = 32; c.b_tile = nb * 128; c.ksteps = 4; c.per_tile = 4; }
                // every operand read must stay inside its region: A region 96 KB, B region 64 KB (tile pitch >= the extent one tile's MMAs touch)
                c.a_tiles = (int)(96u * 1024u / c.a_tile); c.b_tiles = (int)(64u * 1024u / c.b_tile);
                if (c.a_tiles < 1 || c.b_tiles < 1) continue;
                double cyc[2] = {0, 0};
                for (int all = 0; all < 2; ++all) {
                    const int grid = all ? sms : 1;
                    CHECK(cudaMemset(d_cycles, 0, sizeof(unsigned long long) * sms));
                    umma_rate_kernel<<<grid, 128, 176 * 1024>>>(c, d_cycles);       // warm-up
                    umma_rate_kernel<<<grid, 128, 176 * 1024>>>(c, d_cycles);
                    CHECK(cudaDeviceSynchronize());
                    unsigned long long h[256] = {0};
                    CHECK(cudaMemcpy(h, d_cycles, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
                    unsigned long long mx = 0;
                    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                    cyc[all] = (double)mx / iters;
                }
                const double bytes = (c.a_in_tmem ? 0.0 : 4096.0) + n * 32.0;
                printf("%-34s %5d %6s %12.1f %12.1f %10.1f %10.1f\n", name, n, tf32 ? "tf32" : "f16", cyc[0], cyc[1], bytes / cyc[0], 128.0 * n / 256.0);
            }
    cudaFree(d_cycles);
    return 0;
}
