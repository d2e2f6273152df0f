// SM partitions for the factorisation's task graph (CUDA green contexts, driver API >= 12.4).
//
// The panel chain of the blocked Cholesky (64x64 factorisations, fused triangular leaves: one to a
// few dozen CTAs each, 74-210 KB of shared memory per CTA) shares the GPU with the bulk trailing
// updates (DMMA GEMM CTAs, 52 KB, four per SM).  Stream priorities do not help it: a retiring GEMM CTA
// frees less shared memory than a chain CTA needs, so the slot goes back to the next GEMM CTA and the
// chain waits (measured: a 14-block diagonal supertile takes 0.66 ms alone and 1.5-1.8 ms next to the
// updates).  A green context gives the chain its own SMs; the updates get the rest.  The driver entry
// points are fetched through cudaGetDriverEntryPoint, so the library still links against cudart only.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "feba_kernels.h"

namespace feba {

namespace {

template <typename F>
bool entry(const char* name, F* fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess ||
        !p) {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

}  // namespace

int green_create(int device, int reserve_sms, GreenPair* out, char* err, size_t errlen) {
    *out = GreenPair{};
    CUresult (*pDeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*pGetRes)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*pSplit)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) =
        nullptr;
    CUresult (*pDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*pCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    if (!entry("cuDeviceGet", &pDeviceGet) || !entry("cuDeviceGetDevResource", &pGetRes) ||
        !entry("cuDevSmResourceSplitByCount", &pSplit) || !entry("cuDevResourceGenerateDesc", &pDesc) ||
        !entry("cuGreenCtxCreate", &pCreate)) {
        snprintf(err, errlen, "green contexts: driver entry points not available");
        return -1;
    }
    CUdevice dev;
    CUresult rc = pDeviceGet(&dev, device);
    CUdevResource total, part, rest;
    if (rc == CUDA_SUCCESS) rc = pGetRes(dev, &total, CU_DEV_RESOURCE_TYPE_SM);
    unsigned int groups = 1;
    if (rc == CUDA_SUCCESS) rc = pSplit(&part, &groups, &total, &rest, 0, (unsigned int)reserve_sms);
    if (rc == CUDA_SUCCESS && (groups != 1 || rest.sm.smCount == 0)) rc = CUDA_ERROR_INVALID_VALUE;
    CUdevResourceDesc dA = nullptr, dB = nullptr;
    if (rc == CUDA_SUCCESS) rc = pDesc(&dA, &part, 1);
    if (rc == CUDA_SUCCESS) rc = pDesc(&dB, &rest, 1);
    CUgreenCtx gA = nullptr, gB = nullptr;
    if (rc == CUDA_SUCCESS) rc = pCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM);
    if (rc == CUDA_SUCCESS) rc = pCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM);
    if (rc != CUDA_SUCCESS) {
        snprintf(err, errlen, "green contexts: driver call failed (CUresult %d)", (int)rc);
        GreenPair tmp;
        tmp.chain = gA;
        tmp.bulk = gB;
        green_destroy(&tmp);
        return -1;
    }
    out->chain = gA;
    out->bulk = gB;
    out->chain_sms = (int)part.sm.smCount;
    out->bulk_sms = (int)rest.sm.smCount;
    return 0;
}

int green_stream(void* gctx, int priority, cudaStream_t* out) {
    CUresult (*pStream)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    if (!entry("cuGreenCtxStreamCreate", &pStream)) return -1;
    CUstream s = nullptr;
    if (pStream(&s, static_cast<CUgreenCtx>(gctx), CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS) return -1;
    *out = s;
    return 0;
}

void green_destroy(GreenPair* g) {
    CUresult (*pDestroy)(CUgreenCtx) = nullptr;
    if (!entry("cuGreenCtxDestroy", &pDestroy)) return;
    if (g->chain) pDestroy(static_cast<CUgreenCtx>(g->chain));
    if (g->bulk) pDestroy(static_cast<CUgreenCtx>(g->bulk));
    g->chain = g->bulk = nullptr;
}

}  // namespace feba
