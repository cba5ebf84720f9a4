// umma_rate - how fast does one SM retire tcgen05.mma when the operands stream from shared memory?  (measurement tool, not product)
//
// The denoiser's level-0 convs are M = 128, N = 32 MMAs whose A operand (4 KB) and B operand (1 KB) are read from shared memory through
// NO-SWIZZLE K-major descriptors (dn_conv_tc.cuh: the halo-tile trick needs 16-byte tap offsets).  ncu on enc1.l2a: ~75 cycles per MMA
// against a 16-cycle tensor floor, with the tensor pipe at 24 % and the shared-memory operand wavefronts at 57 % - neither saturated.
// This probe separates the candidate causes, per SM and with nothing else running:
//   1. cycles per MMA for N = 16 .. 256 with the conv engine's descriptors (SBO 160 B, LBO 2880 B, 9 taps x 2 K steps per staged tile),
//      with dense no-swizzle core matrices, with SWIZZLE_128B operands, and with the A operand in TMEM;
//   2. the same with the chain of MMAs spread round-robin over 1 / 2 / 4 / 8 ACCUMULATORS: if consecutive MMAs into one accumulator
//      serialise on the accumulate (a fixed latency per dependent MMA), independent accumulators pipeline and the rate rises;
//   3. M = 128 per CTA only (cta_group::1).
// Round 1's version issued from divergent code (`if (elected) for (...) mma`), which costs ~100+ cycles per MMA in R2UR / ELECT
// sequences (profiles/r3a_umma_rate_divergent_issue.txt: every configuration ~105-140 cycles regardless of N).  Here the issue loop is
// warp-uniform with elect.sync inside, unrolled 18 MMAs per trip, descriptors as 32-bit halves - the conv engine's issue pattern.
// Numerics are irrelevant (shared memory is zero filled); only descriptor validity and timing matter.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/umma_rate tools/microbench/umma_rate.cu
//   run:   tools/microbench/umma_rate [trips=128]      (one line per configuration; run it under `timeout`)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)

#define PER_TRIP 18

struct Cfg {
    int n;                    // MMA N
    int layout;               // UMMA layout type: 0 = no swizzle, 2 = SWIZZLE_128B
    uint32_t a_lbo, a_sbo;    // bytes
    uint32_t b_lbo, b_sbo;
    uint32_t a_kstep, b_kstep;// bytes the start address advances for the second K step of a tap
    uint32_t a_tap, b_tap;    // bytes the start address advances per tap (pair of MMAs)
    uint32_t a_tile, b_tile;  // bytes between the staged tiles the trips rotate through
    int a_tiles, b_tiles;
    int a_in_tmem;
    int tf32;                 // kind::tf32 (K = 8) instead of kind::f16 (K = 16); operand bytes per MMA are the same
    int trips;                // PER_TRIP MMAs each
    int nacc;                 // accumulators used round-robin (per MMA)
    int fill;                 // 0: shared memory zero filled; 1: pseudo-random finite operands (does the rate depend on the data?)
    int proto;                // 1: the conv engine's stage protocol - warp 2 waits empty[s] and arrives full[s] (8 stages, no TMA), the MMA warp waits
                              //    full[s], issues the trip's 18 MMAs and commits to empty[s]; 2: the same, and warp 3 waits for a per-tile commit (every 2 trips)
    int commit;               // 1: tcgen05.commit to an mbarrier after every trip of 18 MMAs (the conv engine frees a smem stage per chunk)
    int bg;                   // background traffic from warps 2-3 while the MMAs run: 0 none, 1 st.shared.v4 stream, 2 tcgen05.ld stream, 3 ld.shared.v4 stream
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <bool TF32, bool A_TMEM>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc, uint32_t a_tmem) {
    if (A_TMEM) {
        if (TF32) asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
                               ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
        else asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                          ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
    } else {
        if (TF32) asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                               ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
        else asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                          ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
    }
}

template <bool TF32, bool A_TMEM, int NACC>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(const Cfg c, unsigned long long* cycles_out, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2;
    __shared__ uint64_t pfull[8], pempty[8], tfull[4], tempty[4];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int done_flag;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) done_flag = 0;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c.fill) {                                                    // finite fp32 in [1, 2) with random mantissas == finite fp16 pairs
            uint32_t h = (uint32_t)i * 2654435761u + 12345u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            v.x = 0x38003800u | (h & 0x07ff07ffu); v.y = 0x38003800u | ((h * 3u) & 0x07ff07ffu);
            v.z = 0x38003800u | ((h * 7u) & 0x07ff07ffu); v.w = 0x38003800u | ((h * 11u) & 0x07ff07ffu);
        }
        reinterpret_cast<uint4*>(smem)[i] = v;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000;" ::"r"(smem_u32(&bar2)));
        for (int i = 0; i < 8; ++i) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&pfull[i]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&pempty[i]))); }
        for (int i = 0; i < 4; ++i) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tfull[i]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tempty[i]))); }
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
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp == 0) {
        const uint32_t fmt = TF32 ? 2u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(c.n >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t lay_hi = (uint32_t)(c.layout & 7) << 29;                  // layout type, descriptor bits [61, 64)
        const uint32_t a_hi = (c.a_sbo >> 4) | (1u << 14) | lay_hi, b_hi = (c.b_sbo >> 4) | (1u << 14) | lay_hi;
        const uint32_t a_lbo = (c.a_lbo >> 4) << 16, b_lbo = (c.b_lbo >> 4) << 16;
        const uint32_t a0 = smem_u32(smem) >> 4, b0 = smem_u32(smem + 96 * 1024) >> 4;
        // accumulator stride in TMEM columns: N columns each (NACC * N <= 256 keeps the A-in-TMEM tile at column 256 free)
        const uint32_t acc_cols = (uint32_t)c.n;
        unsigned long long t0 = clock64();
        int at = 0, bt = 0;
        for (int trip = 0; trip < c.trips; ++trip) {
            const uint32_t a_base = (a0 + (uint32_t)at * (c.a_tile >> 4)) | a_lbo;
            const uint32_t b_base = (b0 + (uint32_t)bt * (c.b_tile >> 4)) | b_lbo;
            const int pst = trip & 7; const uint32_t pph = (uint32_t)(trip >> 3) & 1u;
            const int tb = (trip >> 1) & 3; const uint32_t tph = (uint32_t)(trip >> 3) & 1u;
            if (c.proto) {
                if (c.proto == 2 && !(trip & 1)) { asm volatile("{\n\t.reg .pred P1;\n\tWE:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DE;\n\tbra WE;\n\tDE:\n\t}" ::"r"(smem_u32(&tempty[tb])), "r"(tph ^ 1u) : "memory"); }
                asm volatile("{\n\t.reg .pred P1;\n\tWF:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DF;\n\tbra WF;\n\tDF:\n\t}" ::"r"(smem_u32(&pfull[pst])), "r"(pph) : "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (elect_one()) {
#pragma unroll
                for (int t = 0; t < PER_TRIP; ++t) {
                    const uint32_t a_lo = a_base + (uint32_t)(t >> 1) * (c.a_tap >> 4) + (uint32_t)(t & 1) * (c.a_kstep >> 4);
                    const uint32_t b_lo = b_base + (uint32_t)(t >> 1) * (c.b_tap >> 4) + (uint32_t)(t & 1) * (c.b_kstep >> 4);
                    const uint32_t d = tmem + (uint32_t)(t % NACC) * acc_cols;
                    mma<TF32, A_TMEM>(d, a_lo, a_hi, b_lo, b_hi, idesc, (trip || t >= NACC) ? 1u : 0u, tmem + 256);
                }
                if (c.commit) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                if (c.proto) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&pempty[pst])) : "memory");
                if (c.proto == 2 && (trip & 1)) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tfull[tb])) : "memory");
            }
            __syncwarp();
            if (++at == c.a_tiles) at = 0;
            if (++bt == c.b_tiles) bt = 0;
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        unsigned long long t1 = clock64();
        if (threadIdx.x == 0) { cycles_out[blockIdx.x] = t1 - t0; done_flag = 1; }
    } else if (warp == 2 && c.proto) {
        for (int trip = 0; trip < c.trips; ++trip) {
            const int pst = trip & 7; const uint32_t pph = (uint32_t)(trip >> 3) & 1u;
            asm volatile("{\n\t.reg .pred P1;\n\tWP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DP;\n\tbra WP;\n\tDP:\n\t}" ::"r"(smem_u32(&pempty[pst])), "r"(pph ^ 1u) : "memory");
            if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&pfull[pst])) : "memory");
            __syncwarp();
        }
    } else if (warp == 3 && c.proto == 2) {
        for (int trip = 1; trip < c.trips; trip += 2) {
            const int tb = (trip >> 1) & 3; const uint32_t tph = (uint32_t)(trip >> 3) & 1u;
            asm volatile("{\n\t.reg .pred P1;\n\tWT2:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DT2;\n\tbra WT2;\n\tDT2:\n\t}" ::"r"(smem_u32(&tfull[tb])), "r"(tph) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[tb])) : "memory");
            __syncwarp();
        }
    } else if (warp == 2 && c.bg >= 4) {
        // the conv engine's producer: halo-tile boxes (c.bg == 4: 10 px x 18 rows x 4 quads, rows of 160 B starting 16 B before a 128-byte
        // boundary; c.bg == 5: the same bytes as 128-byte-aligned rows) streamed into 4 scratch buffers at smem + 120 KB
        __shared__ uint64_t tbar[4];
        if ((threadIdx.x & 31) == 0) {
            for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar[i])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            uint32_t ph[4] = {0, 0, 0, 0};
            int issued = 0, x = 0, y = 0;
            for (int i = 0; i < 4; ++i) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar[i])), "r"(11520u) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(smem + 120 * 1024 + i * 11520)), "l"(&tmap), "r"(smem_u32(&tbar[i])), "r"(c.bg == 4 ? x * 32 - 4 : x * 32), "r"(y), "r"(0) : "memory");
                x = (x + 1) % 30; y = (y + 16) % 400; ++issued;
            }
            while (!done_flag) {
                const int i = issued & 3;
                asm volatile("{\n\t.reg .pred P1;\n\tWT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DT;\n\tbra WT;\n\tDT:\n\t}" ::"r"(smem_u32(&tbar[i])), "r"(ph[i]) : "memory");
                ph[i] ^= 1;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar[i])), "r"(11520u) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(smem + 120 * 1024 + i * 11520)), "l"(&tmap), "r"(smem_u32(&tbar[i])), "r"(c.bg == 4 ? x * 32 - 4 : x * 32), "r"(y), "r"(0) : "memory");
                x = (x + 1) % 30; y = (y + 16) % 400; ++issued;
            }
            for (int k = 0; k < 4; ++k) {                                   // drain before the CTA exits
                const int i = (issued + k) & 3;
                asm volatile("{\n\t.reg .pred P1;\n\tWU:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DU;\n\tbra WU;\n\tDU:\n\t}" ::"r"(smem_u32(&tbar[i])), "r"(ph[i]) : "memory");
            }
            cycles_out[gridDim.x + blockIdx.x] = (unsigned long long)issued;
        }
        __syncwarp();
    } else if (warp >= 2 && c.bg && c.bg < 4) {
        // background traffic next to the MMA stream, in a scratch region past the operands (smem + 164 KB) / TMEM columns 384..
        uint8_t* scratch = smem + 164 * 1024 + (warp - 2) * 4096;
        const int lane = threadIdx.x & 31;
        uint32_t sink = 0;
        unsigned long long iters_done = 0;
        while (!done_flag) {
            ++iters_done;
            if (c.bg == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) reinterpret_cast<uint4*>(scratch)[k * 32 + lane] = make_uint4(k, lane, sink, 1);
            } else if (c.bg == 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { uint32_t x, y, z, w; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(smem_u32(scratch) + (uint32_t)(k * 32 + lane) * 16u)); sink += x + w; }
            } else {
                uint32_t r[16];
                const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                sink += r[0] + r[15];
            }
        }
        if (sink == 0x12345678u) cycles_out[blockIdx.x + 1] = sink;
        if (warp == 2 && (threadIdx.x & 31) == 0) cycles_out[gridDim.x + blockIdx.x] = iters_done;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

typedef void (*KernelFn)(const Cfg, unsigned long long*, const CUtensorMap);
template <bool TF32, bool A_TMEM>
static KernelFn pick_acc(int nacc) {
    switch (nacc) {
        case 1: return umma_rate_kernel<TF32, A_TMEM, 1>;
        case 2: return umma_rate_kernel<TF32, A_TMEM, 2>;
        case 3: return umma_rate_kernel<TF32, A_TMEM, 3>;
        case 6: return umma_rate_kernel<TF32, A_TMEM, 6>;
        default: return umma_rate_kernel<TF32, A_TMEM, 9>;
    }
}
static KernelFn pick(int tf32, int a_tmem, int nacc) {
    if (tf32) return a_tmem ? pick_acc<true, true>(nacc) : pick_acc<true, false>(nacc);
    return a_tmem ? pick_acc<false, true>(nacc) : pick_acc<false, false>(nacc);
}

int main(int argc, char** argv) {
    const int trips = argc > 1 ? atoi(argv[1]) : 128;
    int dev = 0, sms = 0;
    CHECK(cudaSetDevice(dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned long long* d_cycles = nullptr;
    CHECK(cudaMalloc(&d_cycles, sizeof(unsigned long long) * (2 * sms + 1)));
    // a CHW4-like tensor for the TMA background stream: [4 quads][540 rows][1024 px][4 floats]
    float* d_act = nullptr;
    const size_t act_floats = (size_t)4 * 540 * 1024 * 4;
    CHECK(cudaMalloc(&d_act, act_floats * 4));
    CHECK(cudaMemset(d_act, 0, act_floats * 4));
    CUtensorMap tmap, tmap_aligned;
    {
        typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fp = nullptr; cudaDriverEntryPointQueryResult q;
        CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
        cuuint64_t dims[3] = {1024 * 4, 540, 4}; cuuint64_t strides[2] = {1024 * 16, (cuuint64_t)540 * 1024 * 16};
        cuuint32_t box[3] = {40, 18, 4}, box2[3] = {32, 90, 1}, es[3] = {1, 1, 1};        // 160-byte rows x 72  vs  128-byte rows x 90: both 11520 B
        if (((PFN)fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
            ((PFN)fp)(&tmap_aligned, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_act, dims, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map encode failed\n"); return 1; }
    }
    const int Ns[] = {16, 32, 64, 112, 128, 256};
    const int ACCS[] = {1, 2, 3, 6, 9};
    printf("# %d SMs, %d MMAs per CTA, M = 128, 32 operand bytes per row per MMA (K = 16 fp16 / 8 tf32); grid = 1 CTA and 1 CTA per SM\n", sms, trips * PER_TRIP);
    printf("%-34s %5s %5s %4s %12s %12s %10s %8s\n", "operands", "N", "kind", "acc", "cyc/MMA(1)", "cyc/MMA(all)", "B/clk(1)", "floor");
    for (int variant = 0; variant < 5; ++variant)
        for (int tf32 = 0; tf32 < 2; ++tf32)
            for (int n : Ns)
                for (int nacc : ACCS) {
                    if (nacc * n > 256) continue;
                    if (variant != 0 && variant != 3 && nacc != 1 && nacc != 3) continue;     // the full accumulator sweep for the conv engine's layouts only
                    if (tf32 == 0 && variant != 0 && variant != 2) continue;
                    Cfg c = {};
                    c.n = n; c.trips = trips; c.tf32 = tf32; c.nacc = nacc;
                    const char* name = "";
                    const uint32_t nb = (uint32_t)n;
                    if (variant == 0) { name = "conv engine: no-swizzle halo tile"; c.layout = 0; c.a_lbo = 2880; c.a_sbo = 160; c.a_kstep = 2 * 2880; c.a_tap = 16; c.a_tile = 11520;
                                        c.b_lbo = nb * 16; c.b_sbo = 128; c.b_kstep = nb * 32; c.b_tap = nb * 64; c.b_tile = nb * 64 * 9; }
                    if (variant == 1) { name = "dense no-swizzle core matrices"; c.layout = 0; c.a_lbo = 128; c.a_sbo = 256; c.a_kstep = 4096; c.a_tap = 0; c.a_tile = 8192;
                                        c.b_lbo = 128; c.b_sbo = 256; c.b_kstep = nb * 32; c.b_tap = 0; c.b_tile = nb * 64; }
                    if (variant == 2) { name = "SWIZZLE_128B rows of 64 fp16"; c.layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_kstep = 32; c.a_tap = 64; c.a_tile = 16384;
                                        c.b_lbo = 16; c.b_sbo = 1024; c.b_kstep = 32; c.b_tap = 64; c.b_tile = nb * 128; }
                    if (variant == 3) { name = "A in TMEM, B no-swizzle (conv)"; c.layout = 0; c.a_in_tmem = 1; c.a_tile = 4096;
                                        c.b_lbo = nb * 16; c.b_sbo = 128; c.b_kstep = nb * 32; c.b_tap = nb * 64; c.b_tile = nb * 64 * 9; }
                    if (variant == 4) { name = "A in TMEM, B SWIZZLE_128B"; c.layout = 2; c.a_in_tmem = 1; c.a_tile = 4096;
                                        c.b_lbo = 16; c.b_sbo = 1024; c.b_kstep = 32; c.b_tap = 64; c.b_tile = nb * 128; }
                    if (variant == 2 || variant == 4) { c.a_tap = 0; c.b_tap = 0; }
                    if ((variant == 0 || variant == 3) && c.b_tile > 48u * 1024u) { c.b_tap = 0; c.b_tile = nb * 64; }   // large N: all taps share one B block          // 4 K steps of one swizzled row would need 4 taps; keep 2 K steps, same bytes
                    // every operand read must stay inside its region: A region 96 KB, B region 64 KB
                    c.a_tiles = (int)(96u * 1024u / (c.a_tile + 4096)); c.b_tiles = (int)(64u * 1024u / (c.b_tile + nb * 64));
                    if (c.a_tiles < 1 || c.b_tiles < 1) continue;
                    KernelFn fn = pick(tf32, c.a_in_tmem, nacc);
                    CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024));
                    double cyc[2] = {0, 0};
                    for (int all = 0; all < 2; ++all) {
                        const int grid = all ? sms : 1;
                        CHECK(cudaMemset(d_cycles, 0, sizeof(unsigned long long) * sms));
                        fn<<<grid, 128, 176 * 1024>>>(c, d_cycles, tmap);       // warm-up
                        fn<<<grid, 128, 176 * 1024>>>(c, d_cycles, tmap);
                        CHECK(cudaDeviceSynchronize());
                        unsigned long long h[256] = {0};
                        CHECK(cudaMemcpy(h, d_cycles, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
                        unsigned long long mx = 0;
                        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                        cyc[all] = (double)mx / (trips * PER_TRIP);
                    }
                    const double bytes = (c.a_in_tmem ? 0.0 : 4096.0) + n * 32.0;
                    printf("%-34s %5d %5s %4d %12.1f %12.1f %10.1f %8.1f\n", name, n, tf32 ? "tf32" : "f16", nacc, cyc[0], cyc[1], bytes / cyc[0], 128.0 * n / 256.0);
                }
    printf("# conv engine layout, N = 32: operand data and background traffic (warps 2-3)\n");
    printf("%-6s %-8s %-22s %12s %12s\n", "kind", "data", "background", "cyc/MMA(1)", "cyc/MMA(all)");
    for (int tf32 = 0; tf32 < 2; ++tf32)
        for (int fill = 0; fill < 2; ++fill)
            for (int bg = 0; bg < 9; ++bg) {
                Cfg c = {};
                if (bg >= 4 && fill) continue;
                const uint32_t nb = 32;
                c.n = 32; c.trips = trips; c.tf32 = tf32; c.nacc = 1; c.fill = fill; c.bg = (bg == 4 || bg >= 7) ? 0 : (bg > 4 ? bg - 1 : bg); c.commit = bg == 4; c.proto = bg >= 7 ? bg - 6 : 0;
                c.layout = 0; c.a_lbo = 2880; c.a_sbo = 160; c.a_kstep = 2 * 2880; c.a_tap = 16; c.a_tile = 11520;
                c.b_lbo = nb * 16; c.b_sbo = 128; c.b_kstep = nb * 32; c.b_tap = nb * 64; c.b_tile = nb * 64 * 9;
                c.a_tiles = (int)(96u * 1024u / (c.a_tile + 4096)); c.b_tiles = 1;                  // B region 96..117 KB, TMA scratch from 120 KB
                KernelFn fn = pick(tf32, 0, 1);
                CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024));
                double cyc[2] = {0, 0};
                unsigned long long tma_issued = 0;
                for (int all = 0; all < 2; ++all) {
                    const int grid = all ? sms : 1;
                    CHECK(cudaMemset(d_cycles, 0, sizeof(unsigned long long) * (2 * sms + 1)));
                    fn<<<grid, 128, 176 * 1024>>>(c, d_cycles, bg == 6 ? tmap_aligned : tmap);
                    fn<<<grid, 128, 176 * 1024>>>(c, d_cycles, bg == 6 ? tmap_aligned : tmap);
                    CHECK(cudaDeviceSynchronize());
                    unsigned long long h[256] = {0};
                    CHECK(cudaMemcpy(h, d_cycles, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
                    unsigned long long mx = 0;
                    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                    cyc[all] = (double)mx / (trips * PER_TRIP);
                    CHECK(cudaMemcpy(&tma_issued, d_cycles + grid, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                }
                const char* bgn[] = {"none", "st.shared.v4 stream", "tcgen05.ld stream", "ld.shared.v4 stream", "commit every 18 MMAs", "TMA 160-byte rows -16 B", "TMA 128-byte rows", "stage protocol (full/empty)", "stage + tile protocol"};
                printf("%-6s %-8s %-24s %12.1f %12.1f   background iterations per 18 MMAs (all): %.2f\n", tf32 ? "tf32" : "f16", fill ? "random" : "zeros", bgn[bg], cyc[0], cyc[1], (bg >= 5 || (bg >= 1 && bg <= 3)) ? (double)tma_issued / trips : 0.0);
            }
    cudaFree(d_cycles);
    return 0;
}
