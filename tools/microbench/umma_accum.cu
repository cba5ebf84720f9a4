// umma_accum - how does tcgen05.mma accumulate?  (measurement tool, not product)
//
// The denoiser's PTD_DN_3XTF32 mode landed at rel-L2 1.0e-5 against the fp32 oracle where a CPU emulation with round-to-nearest fp32
// accumulation predicted 1.5e-6.  This probe feeds ONE accumulator tile (M = 128, N = 16) a chain of `reps` MMAs whose
// operands are exactly representable (tf32 / fp16 values), so every product is exact in fp32 and the only error is the accumulation,
// and compares the result with (a) the exact sum in double, (b) an fp32 round-to-nearest chain (one add per MMA of the exact K-step dot
// product), (c) the same chain with round-toward-zero.  Output: signed mean / rms / max error in units of the result's fp32 ulp, and the
// fraction of the 2048 accumulators that match each model bit for bit.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/umma_accum tools/microbench/umma_accum.cu
//   run:   tools/microbench/umma_accum
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)

#define T_TILES 16            // distinct operand tiles in shared memory; MMA r uses tile r % T_TILES
#define A_TILE 4096           // 128 rows x 32 B
#define B_TILE 512            // 16 rows x 32 B
#define NCOL 16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// a, b: operand tiles already in the dense no-swizzle K-major core-matrix layout (element (row, 16-byte k vector v) at (row / 8) * 256 + v * 128 + (row % 8) * 16)
__global__ void __launch_bounds__(128, 1) umma_accum_kernel(const uint8_t* a, const uint8_t* b, int reps, int tf32, float* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;
    uint8_t* sb = smem + T_TILES * A_TILE;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < T_TILES * A_TILE / 16; i += blockDim.x) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a)[i];
    for (int i = threadIdx.x; i < T_TILES * B_TILE / 16; i += blockDim.x) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b)[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t fmt = tf32 ? 2u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
        if (pred) {
            for (int i = 0; i < reps; ++i) {
                const int t = i % T_TILES;
                const uint64_t ad = make_desc(smem_u32(sa + t * A_TILE), 128, 256);
                const uint64_t bd = make_desc(smem_u32(sb + t * B_TILE), 128, 256);
                if (tf32) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1 : 0) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1 : 0) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
    }
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
    }
}

static float round_tf32(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x1000u; u &= 0xffffe000u;
    float r; memcpy(&r, &u, 4); return r;
}
static float add_rz(float a, float b) {                 // fp32 add rounded toward zero, through double (exact for one add of two floats? no: emulate)
    const double s = (double)a + (double)b;             // exact in double unless exponents differ by > 29 bits; fine for these magnitudes
    float r = (float)s;                                 // RN
    if (std::fabs((double)r) > std::fabs(s)) r = std::nextafterf(r, 0.0f);
    return r;
}
static double urand(uint64_t& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(s >> 11) / 9007199254740992.0; }

int main() {
    CHECK(cudaSetDevice(0));
    CHECK(cudaFuncSetAttribute(umma_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    uint8_t *d_a, *d_b; float* d_out;
    CHECK(cudaMalloc(&d_a, T_TILES * A_TILE)); CHECK(cudaMalloc(&d_b, T_TILES * B_TILE)); CHECK(cudaMalloc(&d_out, 128 * 16 * 4));
    printf("# tcgen05.mma accumulation probe: M = 128, N = 16, one accumulator, `reps` chained MMAs; operands exactly representable\n");
    printf("%-6s %-9s %5s | %28s | %28s | %28s | %9s %9s\n", "kind", "data", "reps", "HW: mean/rms/max err [ulp]", "RN chain: mean/rms/max", "RZ chain: mean/rms/max", "HW==RN", "HW==RZ");
    for (int tf32 = 1; tf32 >= 0; --tf32)
        for (int data = 0; data < 2; ++data) {
            const int K = tf32 ? 8 : 16;                 // elements per 32-byte row
            std::vector<float> A((size_t)T_TILES * 128 * K), B((size_t)T_TILES * NCOL * K);
            uint64_t seed = 1234 + data * 77 + tf32;
            auto gen = [&](float& v) {
                double u = urand(seed);
                double x = data == 0 ? 0.5 + 0.5 * u : (u - 0.5) * 2.0;           // positive [0.5, 1)  /  signed (-1, 1)
                float f = (float)x;
                v = tf32 ? round_tf32(f) : __half2float(__float2half_rn(f));
            };
            for (auto& v : A) gen(v);
            for (auto& v : B) gen(v);
            std::vector<uint8_t> pa(T_TILES * A_TILE), pb(T_TILES * B_TILE);
            const int epv = K / 2;                       // elements per 16-byte vector: 4 tf32 / 8 fp16
            auto put = [&](std::vector<uint8_t>& dst, size_t tile_off, int row, int k, float v) {
                const size_t off = tile_off + (size_t)(row / 8) * 256 + (size_t)(k / epv) * 128 + (size_t)(row % 8) * 16 + (size_t)(k % epv) * (tf32 ? 4 : 2);
                if (tf32) memcpy(&dst[off], &v, 4);
                else { __half h = __float2half_rn(v); memcpy(&dst[off], &h, 2); }
            };
            for (int t = 0; t < T_TILES; ++t) {
                for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) put(pa, (size_t)t * A_TILE, m, k, A[((size_t)t * 128 + m) * K + k]);
                for (int n = 0; n < NCOL; ++n) for (int k = 0; k < K; ++k) put(pb, (size_t)t * B_TILE, n, k, B[((size_t)t * NCOL + n) * K + k]);
            }
            CHECK(cudaMemcpy(d_a, pa.data(), pa.size(), cudaMemcpyHostToDevice));
            CHECK(cudaMemcpy(d_b, pb.data(), pb.size(), cudaMemcpyHostToDevice));
            const int REPS[] = {1, 2, 4, 18, 72, 216, 1024};
            for (int reps : REPS) {
                umma_accum_kernel<<<1, 128, 96 * 1024>>>(d_a, d_b, reps, tf32, d_out);
                CHECK(cudaDeviceSynchronize());
                std::vector<float> out(128 * 16);
                CHECK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
                double st[3][3] = {{0}};                 // [hw, rn, rz][mean, sumsq, max]
                int eq_rn = 0, eq_rz = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < NCOL; ++n) {
                        double exact = 0; float rn = 0.f, rz = 0.f;
                        for (int r = 0; r < reps; ++r) {
                            const int t = r % T_TILES;
                            double dot = 0;
                            for (int k = 0; k < K; ++k) dot += (double)A[((size_t)t * 128 + m) * K + k] * (double)B[((size_t)t * NCOL + n) * K + k];
                            exact += dot;
                            // chain models: the K-step dot product exact (it is, in double), then ONE fp32 add per MMA
                            const double srn = (double)rn + dot; rn = (float)srn;
                            const double srz = (double)rz + dot; float q = (float)srz; if (std::fabs((double)q) > std::fabs(srz)) q = std::nextafterf(q, 0.0f); rz = q;
                        }
                        const float ref = (float)exact;
                        const double ulp = std::ldexp(1.0, std::ilogb(std::fabs(ref) > 0 ? ref : 1.0f) - 23);
                        const float v[3] = {out[m * 16 + n], rn, rz};
                        for (int i = 0; i < 3; ++i) {
                            const double e = ((double)v[i] - exact) / ulp;
                            st[i][0] += e; st[i][1] += e * e; if (std::fabs(e) > st[i][2]) st[i][2] = std::fabs(e);
                        }
                        eq_rn += out[m * 16 + n] == rn; eq_rz += out[m * 16 + n] == rz;
                    }
                const double cnt = 128.0 * NCOL;
                printf("%-6s %-9s %5d |", tf32 ? "tf32" : "f16", data == 0 ? "positive" : "signed", reps);
                for (int i = 0; i < 3; ++i) printf(" %+9.2f %8.2f %8.2f |", st[i][0] / cnt, std::sqrt(st[i][1] / cnt), st[i][2]);
                printf(" %8.1f%% %8.1f%%\n", 100.0 * eq_rn / cnt, 100.0 * eq_rz / cnt);
            }
        }
    (void)add_rz;
    return 0;
}
