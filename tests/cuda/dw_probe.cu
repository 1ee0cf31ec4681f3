// Standalone probe for k_umma_dw (MN-major tcgen05 GEMM): C = Aᵀ·B on small exact-integer inputs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I go-ctr_b200/csrc -o gpurun_out/dw_probe tests/cuda/dw_probe.cu
#include "umma_gemm.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace ctr;
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN enc;
static void mk(CUtensorMap* m, float* p, uint64_t rows, uint64_t cols, uint64_t ld) {
    cuuint64_t gd[2] = {cols, rows}; cuuint64_t gs[1] = {ld * 4}; cuuint32_t box[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d\n", (int)r); exit(1); }
}
int main(int argc, char** argv) {
    int K = argc > 1 ? atoi(argv[1]) : 64, Mw = 256, Nw = 224;
    cudaDriverEntryPointQueryResult q; void* fp = nullptr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q); enc = (PFN)fp;
    std::vector<float> A((size_t)K * Mw), B((size_t)K * Nw), C((size_t)Mw * Nw, 0.f), R((size_t)Mw * Nw, 0.f);
    for (int k = 0; k < K; k++) for (int m = 0; m < Mw; m++) A[(size_t)k * Mw + m] = (float)(((k * 7 + m * 3) % 11) - 5);
    for (int k = 0; k < K; k++) for (int n = 0; n < Nw; n++) B[(size_t)k * Nw + n] = (float)(((k * 5 + n) % 7) - 3);
    for (int m = 0; m < Mw; m++) for (int n = 0; n < Nw; n++) { double s = 0; for (int k = 0; k < K; k++) s += (double)A[(size_t)k * Mw + m] * B[(size_t)k * Nw + n]; R[(size_t)m * Nw + n] = (float)s; }
    float *dA, *dB, *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0, C.size() * 4);
    CUtensorMap mA, mB; mk(&mA, dA, K, Mw, Mw); mk(&mB, dB, K, Nw, Nw);
    umma::DwArgs a{}; a.K = K; a.na = 8; a.nb = 7; a.M = Mw; a.N = Nw; a.C = dC; a.ldc = Nw; a.stages = 3; a.dbg = argc > 3 ? atoi(argv[3]) : 0;
    size_t smem = (size_t)a.stages * (8 * 4096 + a.nb * 4096) + 8 * (3 * a.stages + 2) + 16 + 1024;
    cudaFuncSetAttribute(umma::k_umma_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    int grid = argc > 2 ? atoi(argv[2]) : 1;
    umma::k_umma_dw<<<grid, 448, smem>>>(mA, mB, a);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < C.size(); i++) { double d = fabs((double)C[i] - R[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) bad++; }
    printf("K=%d grid=%d max|err|=%g bad=%d/%zu\n", K, grid, maxerr, bad, C.size());
    for (int m : {0, 1, 5, 31, 32, 127, 128, 200, 255}) { printf("m=%3d:", m); for (int n : {0, 1, 4, 31, 32, 100, 223}) printf("  %g/%g", C[(size_t)m * Nw + n], R[(size_t)m * Nw + n]); printf("\n"); }
    return 0;
}
