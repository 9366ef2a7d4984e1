// tma_probe.cu -- development probe: one TMA box load of an fp64 field through the helpers of
// dc_stage3.h.  nvcc -gencode arch=compute_100a,code=sm_100a -I climate_model_b200/csrc ...
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "dc_stage3.h"
using namespace dc;

struct Probe {
    int mode;
    unsigned bytes;
    int style;
    int rank;
    TmaMap m;
    const TmaMap *gm;   // descriptor in global memory (or NULL: use the kernel parameter)
    double *out;
    const double *src;
};

__global__ void k_probe(const __grid_constant__ Probe p, int x, int y, int z)
{
    extern __shared__ unsigned char raw[];
    const unsigned a = (unsigned)__cvta_generic_to_shared(raw);
    unsigned char *base = raw + ((128u - (a & 127u)) & 127u);
    double *dst = reinterpret_cast<double *>(base);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(base + S3_PL * 8);
    const int tid = threadIdx.x;
    if (tid == 0) {
        s3_mbar_init(bar, 1);
        s3_mbar_init_fence();
    }
    __syncthreads();
    if (p.style == 1 && tid < 32) {
        if (tid == 0) s3_mbar_expect(bar, p.bytes);
        __syncwarp();
        const TmaMap *mp = p.gm ? p.gm : &p.m;
        asm volatile(
            "{\n\t.reg .pred q;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];\n\t}" ::"r"(s3_smem_u32(dst)),
            "l"(reinterpret_cast<unsigned long long>(mp)), "r"(s3_smem_u32(bar)), "r"(x), "r"(y), "r"(z)
            : "memory");
    }
    if (p.style == 0 && p.mode >= 1 && tid == 0) {
        s3_mbar_expect(bar, p.bytes);
        if (p.mode >= 2 && p.mode < 5 && p.rank == 3) s3_tma_load(dst, p.gm ? p.gm : &p.m, x, y, z, bar);
        if (p.mode >= 2 && p.mode < 5 && p.rank == 2)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4}], [%2];" ::"r"(s3_smem_u32(dst)),
                "l"(reinterpret_cast<unsigned long long>(p.gm ? p.gm : &p.m)), "r"(s3_smem_u32(bar)), "r"(x), "r"(y)
                : "memory");
        if (p.mode == 5)   // plain bulk copy, no tensor map
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(s3_smem_u32(dst)), "l"(p.src), "r"(p.bytes), "r"(s3_smem_u32(bar)) : "memory");
    }
    if (p.mode == 2) s3_mbar_wait(bar, 0);
    if (p.mode == 3) {   // TMA issued, no mbarrier wait: just give it time
        for (int n = 0; n < 2000; n++) __nanosleep(1000);
    }
    if (p.mode == 4 || p.mode == 5) {   // wait without labels inside the asm
        unsigned ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(s3_smem_u32(bar)), "r"(0u)
                : "memory");
        }
    }
    __syncthreads();
    for (int n = tid; n < S3_SN; n += blockDim.x) p.out[n] = p.mode >= 2 ? dst[n] : 1.0;
}

int main(int argc, char **argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 2;
    const int NI = 64, NJ = 20, NK = 4;
    double *h = (double *)malloc(sizeof(double) * NI * NJ * NK), *d, *out;
    for (int n = 0; n < NI * NJ * NK; n++) h[n] = n;
    cudaMalloc(&d, sizeof(double) * NI * NJ * NK);
    cudaMalloc(&out, sizeof(double) * S3_SN);
    cudaMemcpy(d, h, sizeof(double) * NI * NJ * NK, cudaMemcpyHostToDevice);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                           CUtensorMapFloatOOBfill);
    Probe p;
    p.mode = mode;
    p.out = out;
    p.src = d;
    const cuuint64_t dims[3] = {NI, NJ, NK};
    const cuuint64_t strides[2] = {NI * 8, (cuuint64_t)NI * NJ * 8};
    cuuint32_t box[3] = {S3_SW, S3_SH, 1};
    if (argc > 2) box[0] = atoi(argv[2]);
    if (argc > 3) box[1] = atoi(argv[3]);
    const cuuint32_t estr[3] = {1, 1, 1};
    const int f32 = argc > 4 ? atoi(argv[4]) : 0;
    cuuint64_t dims32[3] = {NI * 2, NJ, NK};
    cuuint32_t box32[3] = {box[0] * 2, box[1], 1};
    CUresult r = ((Fn)fp)(&p.m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, f32 ? dims32 : dims, strides, f32 ? box32 : box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    p.bytes = box[0] * box[1] * 8;
    const int rank = argc > 6 ? atoi(argv[6]) : 3;
    if (rank == 2) {
        r = ((Fn)fp)(&p.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    p.rank = rank;
    p.style = argc > 7 ? atoi(argv[7]) : 0;
    const int swz = argc > 8 ? atoi(argv[8]) : 0;
    for (int n = 0; n < 16; n++) printf("%016llx%c", (unsigned long long)p.m.opaque[n], n % 4 == 3 ? '\n' : ' ');
    printf("encode rc=%d entry=%p q=%d\n", (int)r, fp, (int)q);
    p.gm = nullptr;
    if (argc > 5 && atoi(argv[5])) {
        TmaMap *g;
        cudaMalloc(&g, sizeof(TmaMap));
        cudaMemcpy(g, &p.m, sizeof(TmaMap), cudaMemcpyHostToDevice);
        p.gm = g;
    }
    const int smem = S3_PL * 8 + 64 + 128;
    const int X = argc > 9 ? atoi(argv[9]) : 3;
    k_probe<<<1, 128, smem>>>(p, X, 2, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        double *o = (double *)malloc(sizeof(double) * S3_SN);
        cudaMemcpy(o, out, sizeof(double) * S3_SN, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int b = 0; b < S3_SH; b++)
            for (int a = 0; a < S3_SW; a++) {
                const double want = mode >= 2 ? (double)((1 * NJ + (2 + b)) * NI + X + a) : 1.0;
                if (o[b * S3_SW + a] != want) bad++;
            }
        printf("mismatches: %d  (o[0]=%g o[37]=%g)\n", bad, o[0], o[37]);
    }
    return 0;
}
