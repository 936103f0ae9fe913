// cvad_dev.cu -> libcvad_b200_dev.so: development / measurement hooks for the tcgen05 building blocks.
// Test infrastructure, built next to the product library but never loaded by it (include/cutter_vad_b200_dev.h).
#include <string>

#include "../../include/cutter_vad_b200.h"
#include "../../include/cutter_vad_b200_dev.h"
#include "cvad_tc_dev.cuh"

namespace {
thread_local std::string g_dev_error;
int dev_fail(int code, const std::string &msg) { g_dev_error = msg; return code; }
#define DEV_TRY(call)                                                                          \
    do {                                                                                       \
        cudaError_t _st = (call);                                                              \
        if (_st != cudaSuccess) return dev_fail(CVAD_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_st)); \
    } while (0)
}  // namespace

extern "C" {

const char *cvad_dev_last_error(void) { return g_dev_error.c_str(); }

// Hardware probe of the tcgen05 path (test hook): D[128][32] = A[128][256] * B[32][256]^T, operands are
// raw bf16 bit patterns on the host, D is float32.  Returns 0, or a negative CVAD_E_* code.
int cvad_tc_probe(int device, const uint16_t *a_bf16, const uint16_t *b_bf16, float *d_out) {
    using namespace cvad::tc;
    if (!a_bf16 || !b_bf16 || !d_out) return CVAD_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return dev_fail(CVAD_E_NOGPU, "cudaSetDevice failed");
    __nv_bfloat16 *dA = nullptr, *dB = nullptr;
    float *dD = nullptr;
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&dA), kProbeM * kProbeK * 2));
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&dB), kProbeN * kProbeK * 2));
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&dD), kProbeM * kProbeN * 4));
    DEV_TRY(cudaMemcpy(dA, a_bf16, kProbeM * kProbeK * 2, cudaMemcpyHostToDevice));
    DEV_TRY(cudaMemcpy(dB, b_bf16, kProbeN * kProbeK * 2, cudaMemcpyHostToDevice));
    DEV_TRY(cudaMemset(dD, 0xFF, kProbeM * kProbeN * 4));
    DEV_TRY(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kProbeSmem));
    tc_probe_kernel<<<1, 128, kProbeSmem>>>(dA, dB, dD);
    DEV_TRY(cudaGetLastError());
    DEV_TRY(cudaDeviceSynchronize());
    DEV_TRY(cudaMemcpy(d_out, dD, kProbeM * kProbeN * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return CVAD_OK;
}


// MN-major B operand probe: a_bf16[128][64], b_bf16[96][64] -> d_out[128][160] (columns 0..95: A * B^T, 96..159: A * B[32..95]^T)
int cvad_tc_probe_mn(int device, const uint16_t *a_bf16, const uint16_t *b_bf16, float *d_out) {
    using namespace cvad::tc;
    if (!a_bf16 || !b_bf16 || !d_out) return CVAD_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return dev_fail(CVAD_E_NOGPU, "cudaSetDevice failed");
    __nv_bfloat16 *dA = nullptr, *dB = nullptr;
    float *dD = nullptr;
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&dA), 128 * 64 * 2));
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&dB), 96 * 64 * 2));
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&dD), 128 * 160 * 4));
    DEV_TRY(cudaMemcpy(dA, a_bf16, 128 * 64 * 2, cudaMemcpyHostToDevice));
    DEV_TRY(cudaMemcpy(dB, b_bf16, 96 * 64 * 2, cudaMemcpyHostToDevice));
    DEV_TRY(cudaMemset(dD, 0xFF, 128 * 160 * 4));
    DEV_TRY(cudaFuncSetAttribute(tc_probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kProbeMnSmem));
    tc_probe_mn_kernel<<<1, 128, kProbeMnSmem>>>(dA, dB, dD);
    DEV_TRY(cudaGetLastError());
    DEV_TRY(cudaDeviceSynchronize());
    DEV_TRY(cudaMemcpy(d_out, dD, 128 * 160 * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return CVAD_OK;
}

int cvad_tc_rate(int device, int M, int N, int reps, int a_tiles, int n_acc, int grid, long long *out2) {
    return cvad_tc_rate_mn(device, M, N, reps, a_tiles, n_acc, grid, 0, out2);
}

int cvad_tc_rate_mn(int device, int M, int N, int reps, int a_tiles, int n_acc, int grid, int b_mn, long long *out2) {
    using namespace cvad::tc;
    if (b_mn && N % 32) return CVAD_E_INVALID;
    if (!out2 || a_tiles < 1 || a_tiles > 10 || (M != 64 && M != 128) || N < 8 || N > 256 || grid < 1 || n_acc < 1 || n_acc * N > 512) return CVAD_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return dev_fail(CVAD_E_NOGPU, "cudaSetDevice failed");
    long long *d = nullptr;
    const size_t smem = (size_t)a_tiles * 16384 + 32768 + 1024 + 64;
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&d), 16));
    DEV_TRY(cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_rate_kernel<<<grid, 128, smem>>>(M, N, reps, a_tiles, n_acc, b_mn, d);
    DEV_TRY(cudaGetLastError());
    DEV_TRY(cudaDeviceSynchronize());
    DEV_TRY(cudaMemcpy(out2, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return CVAD_OK;
}

int cvad_bulk_rate(int device, int tiles, int depth, int tile_bytes, int grid, size_t src_bytes, long long *out2) {
    using namespace cvad::tc;
    if (!out2 || tiles < 1 || depth < 1 || depth * (size_t)tile_bytes > 200 * 1024 || tile_bytes % 16 || grid < 1 ||
        src_bytes < (size_t)tile_bytes)
        return CVAD_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return dev_fail(CVAD_E_NOGPU, "cudaSetDevice failed");
    long long *d = nullptr;
    unsigned char *src = nullptr;
    const size_t smem = (size_t)depth * tile_bytes + 1024 + 8 * depth + 64;
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&d), 16));
    DEV_TRY(cudaMalloc(reinterpret_cast<void **>(&src), src_bytes));
    DEV_TRY(cudaMemset(src, 1, src_bytes));
    DEV_TRY(cudaFuncSetAttribute(bulk_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) {   // second run: source resident in L2
        bulk_rate_kernel<<<grid, 64, smem>>>(src, src_bytes, tiles, depth, tile_bytes, d);
        DEV_TRY(cudaGetLastError());
        DEV_TRY(cudaDeviceSynchronize());
    }
    DEV_TRY(cudaMemcpy(out2, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(src);
    return CVAD_OK;
}


}  // extern "C"
