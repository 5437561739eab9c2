// minimal TMA 2-D tile load test (debug aid): tma_min <elem_bytes 1|4> <boxW> <boxH> <swizzle 0|3> <l2promo 0|2>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int bytes, int c0, int r0, unsigned *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned long long *barp = reinterpret_cast<unsigned long long *>(smem + bytes);
    const unsigned barA = (unsigned)__cvta_generic_to_shared(barp);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barA), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&tm), "r"(c0), "r"(r0), "r"(barA) : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(barA), "r"(0) : "memory");
    unsigned s = 0;
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) s += smem[i];
    atomicAdd(out, s);
}
int main(int argc, char **argv) {
    const int es = argc > 1 ? atoi(argv[1]) : 1, bw = argc > 2 ? atoi(argv[2]) : 128, bh = argc > 3 ? atoi(argv[3]) : 128;
    const int sw = argc > 4 ? atoi(argv[4]) : 0, l2 = argc > 5 ? atoi(argv[5]) : 2;
    const int W = 1616, H = 1601;
    std::vector<unsigned char> h((size_t)W * H * es);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W * es};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, est[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, est,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw, (CUtensorMapL2promotion)l2,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int bytes = bw * bh * es;
    unsigned *out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 64);
    const int c0 = 704, r0 = 800;
    k<<<1, 256, bytes + 64>>>(tm, bytes, c0, r0, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned got = 0; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
    unsigned ref = 0;
    for (int rr = 0; rr < bh; rr++) for (int b = 0; b < bw * es; b++) ref += h[((size_t)(r0 + rr) * W + c0) * es + b];
    printf("es=%d box=%dx%d sw=%d l2=%d encode=%d: %s got %u ref %u %s\n", es, bw, bh, sw, l2, (int)r, cudaGetErrorString(e), got, ref,
           (got == ref) ? "MATCH" : "MISMATCH");
    return 0;
}
