// partition.cu — SM partitioning for compute/collective overlap (multi-GPU training).
//
// The persistent training grid fills every SM, so a collective kernel (NCCL all-reduce of the item-table
// deltas) launched beside it cannot start until the grid drains.  eb_partition_streams_create carves the
// device into a green context holding all but `reserve_sms` SMs and returns streams bound to it: kernels
// launched on those streams only ever occupy the partition, the SMs left out stay free for the collective
// running on an ordinary stream.  Driver entry points are resolved through the runtime
// (cudaGetDriverEntryPoint) so the library has no link-time dependency on libcuda.
// The reference has no counterpart (single device).
#include <cuda.h>

#include "common.cuh"

namespace eb {
namespace {

template <typename F>
int drv(const char *name, F *out) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
        return set_err(EB_ERR_CUDA, "driver entry point %s unavailable (%s)", name, cudaGetErrorString(e));
    *out = reinterpret_cast<F>(p);
    return EB_OK;
}

#define EB_DRV(call)                                                                         \
    do {                                                                                     \
        CUresult r_ = (call);                                                                \
        if (r_ != CUDA_SUCCESS) return set_err(EB_ERR_CUDA, "%s failed: CUresult %d", #call, (int)r_); \
    } while (0)

}  // namespace
}  // namespace eb

extern "C" int eb_partition_streams_create(int reserve_sms, int n_streams, void **streams, int *granted_sms) {
    using namespace eb;
    EB_ARG(streams && n_streams >= 1 && n_streams <= 16 && reserve_sms >= 1, "bad argument");
    int dev = 0;
    EB_CUDA(cudaGetDevice(&dev));
    EB_CUDA(cudaFree(0));                                      // primary context must be live
    const int total = sm_count();
    EB_ARG(reserve_sms < total, "reserve_sms=%d must be below the SM count %d", reserve_sms, total);

    CUresult (*p_cuDeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*p_getRes)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*p_split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
    CUresult (*p_desc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
    CUresult (*p_create)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*p_ctxRes)(CUgreenCtx, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*p_stream)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
    int rc;
    if ((rc = drv("cuDeviceGet", &p_cuDeviceGet))) return rc;
    if ((rc = drv("cuDeviceGetDevResource", &p_getRes))) return rc;
    if ((rc = drv("cuDevSmResourceSplitByCount", &p_split))) return rc;
    if ((rc = drv("cuDevResourceGenerateDesc", &p_desc))) return rc;
    if ((rc = drv("cuGreenCtxCreate", &p_create))) return rc;
    if ((rc = drv("cuGreenCtxGetDevResource", &p_ctxRes))) return rc;
    if ((rc = drv("cuGreenCtxStreamCreate", &p_stream))) return rc;

    CUdevice cu_dev;
    EB_DRV(p_cuDeviceGet(&cu_dev, dev));
    CUdevResource all, part, rest;
    EB_DRV(p_getRes(cu_dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    // the driver rounds the group UP to its granularity (8 SMs on sm_90+): ask for the largest multiple of 8
    // that still leaves at least reserve_sms out
    unsigned int want = (unsigned)(total - reserve_sms);
    want -= want % 8;
    EB_ARG(want >= 8, "partition would be empty");
    unsigned int groups = 1;
    EB_DRV(p_split(&part, &groups, &all, &rest, 0, want));
    EB_ARG(groups == 1, "SM split produced %u groups", groups);
    CUdevResourceDesc desc;
    EB_DRV(p_desc(&desc, &part, 1));
    CUgreenCtx gctx;                                           // lives until process exit (streams reference it)
    EB_DRV(p_create(&gctx, desc, cu_dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUdevResource got;
    EB_DRV(p_ctxRes(gctx, &got, CU_DEV_RESOURCE_TYPE_SM));
    for (int s = 0; s < n_streams; ++s) {
        CUstream st;
        EB_DRV(p_stream(&st, gctx, CU_STREAM_NON_BLOCKING, 0));
        streams[s] = (void *)st;
    }
    if (granted_sms) *granted_sms = (int)got.sm.smCount;
    return EB_OK;
}
