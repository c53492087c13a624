// Stand-alone probe of the tcgen05 tf32 building blocks used by render_tc_kernel:
// canonical un-swizzled layouts, descriptor field roles, TMEM read-back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../gaussianformer_b200/csrc/common.cuh"
using namespace gf;

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    return d;
}

// mode bit0: swap A lbo/sbo; bit1: swap B lbo/sbo; bit2: B K-major layout instead of MN-major
__global__ void __launch_bounds__(128) probe(int mode, float *out) {
    __shared__ __align__(128) uint32_t a[128 * 16];
    __shared__ __align__(128) uint32_t b[32 * 16];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool bk = mode & 4;
    // A[m][k] = (m % 7 + 1) + k ; K-major canonical: (m%8)*16 + (m/8)*128 + (k/4)*2048 + (k%4)*4
    for (int k = 0; k < 16; ++k) {
        float v = (float)(tid % 7 + 1) + (float)k;
        uint32_t off = (tid & 7) * 16 + (tid >> 3) * 128 + (k >> 2) * 2048 + (k & 3) * 4;
        a[off / 4] = __float_as_uint(v);
    }
    // B[k][n] = (k % 3 + 1) * (n + 1)
    for (int i = tid; i < 32 * 16; i += 128) {
        int k = i / 32, n = i % 32;
        float v = (float)((k % 3 + 1) * (n + 1));
        uint32_t off;
        if (!bk) off = (n & 3) * 4 + (n >> 2) * 256 + (k & 7) * 16 + (k >> 3) * 128;       // MN-major
        else     off = (n & 7) * 16 + (n >> 3) * 128 + (k >> 2) * 512 + (k & 3) * 4;         // K-major (N=32: 4 groups)
        b[off / 4] = __float_as_uint(v);
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tbase;
    if (mode & 8) {   // prefill D with tid*100 + col, then accumulate on top
        uint32_t w[32];
        for (int i = 0; i < 32; ++i) w[i] = __float_as_uint((float)(tid * 100 + i));
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
            "tcgen05.wait::st.sync.aligned;\n"
            :: "r"(tmem + ((uint32_t)(warp * 32) << 16)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
               "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]),
               "r"(w[18]), "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]),
               "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31]) : "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0 && !(mode & 16)) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((bk ? 0u : 1u) << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t albo = 2048, asbo = 128;
            uint32_t blbo = bk ? 512 : 128, bsbo = bk ? 128 : 256;
            uint64_t da = (mode & 1) ? mkdesc(smem_u32(a) + ks * 4096, asbo, albo) : mkdesc(smem_u32(a) + ks * 4096, albo, asbo);
            uint32_t bstart = smem_u32(b) + (bk ? ks * 1024 : ks * 128);
            uint64_t db = (mode & 2) ? mkdesc(bstart, bsbo, blbo) : mkdesc(bstart, blbo, bsbo);
            uint32_t acc = (ks > 0) || (mode & 8);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (tid == 0 && (mode & 16)) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
    for (int i = 0; i < 32; ++i) out[tid * 32 + i] = __uint_as_float(r[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
    float *d;
    cudaMalloc(&d, 128 * 32 * 4);
    static float h[128 * 32];
    const int modes[] = {0, 4, 2, 6, 8, 12, 24};
    for (int mi = 0; mi < 7; ++mi) {
        int mode = modes[mi];
        cudaMemset(d, 0, 128 * 32 * 4);
        probe<<<1, 128>>>(mode, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) {
                double ref = 0;
                for (int k = 0; k < 16; ++k) ref += ((m % 7 + 1) + k) * (double)((k % 3 + 1) * (n + 1));
                double err = fabs(h[m * 32 + n] - ref);
                if (err > 1e-3 * fabs(ref)) bad++;
                if (err > maxerr) maxerr = err;
            }
        printf("mode %2d (Bswap=%d Bkmajor=%d prefill=%d nomma=%d): bad=%d/4096 maxerr=%g  D[0][0..3]=%g %g %g %g  D[9][0..2]=%g %g %g D[127][31]=%g\n", mode,
               (mode >> 1) & 1, (mode >> 2) & 1, (mode>>3)&1, (mode>>4)&1, bad, maxerr, h[0], h[1], h[2], h[3], h[9 * 32], h[9*32+1], h[9*32+2], h[127*32+31]);
    }
    double ref00 = 0; for (int k = 0; k < 16; ++k) ref00 += (1 + k) * (double)(k % 3 + 1);
    printf("ref D[0][0]=%g D[0][1]=%g\n", ref00, 2 * ref00);
    return 0;
}
