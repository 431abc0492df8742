// rn_cuda_api.hpp -- the CUDA driver API, loaded with dlopen at first use so that librainier_cuda.so can be
// loaded (and its emitter / NVRTC path exercised) on a box without a GPU driver.  There is no CPU fallback:
// every entry point that needs a device fails with RN_E_CUDA when the driver or a device is missing.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace rn {
namespace cu {

typedef int CUresult;
typedef int CUdevice;
typedef struct CUctx_st* CUcontext;
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
typedef struct CUstream_st* CUstream;
typedef struct CUevent_st* CUevent;
typedef unsigned long long CUdeviceptr;

struct CUDA_MEMCPY2D {
  size_t srcXInBytes, srcY;
  int srcMemoryType;
  const void* srcHost;
  CUdeviceptr srcDevice;
  void* srcArray;
  size_t srcPitch;
  size_t dstXInBytes, dstY;
  int dstMemoryType;
  void* dstHost;
  CUdeviceptr dstDevice;
  void* dstArray;
  size_t dstPitch;
  size_t WidthInBytes, Height;
};
enum { CU_MEMORYTYPE_HOST = 1, CU_MEMORYTYPE_DEVICE = 2 };

struct Api {
  CUresult (*cuInit)(unsigned);
  CUresult (*cuDeviceGet)(CUdevice*, int);
  CUresult (*cuDeviceGetCount)(int*);
  CUresult (*cuDeviceGetAttribute)(int*, int, CUdevice);
  CUresult (*cuDevicePrimaryCtxRetain)(CUcontext*, CUdevice);
  CUresult (*cuDevicePrimaryCtxRelease)(CUdevice);
  CUresult (*cuCtxSetCurrent)(CUcontext);
  CUresult (*cuCtxGetCurrent)(CUcontext*);
  CUresult (*cuModuleLoadData)(CUmodule*, const void*);
  CUresult (*cuModuleUnload)(CUmodule);
  CUresult (*cuModuleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*cuMemAlloc)(CUdeviceptr*, size_t);
  CUresult (*cuMemFree)(CUdeviceptr);
  CUresult (*cuMemAllocHost)(void**, size_t);
  CUresult (*cuMemFreeHost)(void*);
  CUresult (*cuMemHostRegister)(void*, size_t, unsigned);
  CUresult (*cuMemHostUnregister)(void*);
  CUresult (*cuPointerGetAttribute)(void*, int, CUdeviceptr);
  CUresult (*cuPointerGetAttributes)(unsigned, int*, void**, CUdeviceptr);
  CUresult (*cuCtxGetDevice)(CUdevice*);
  CUresult (*cuDeviceGetPCIBusId)(char*, int, CUdevice);
  CUresult (*cuMemcpyHtoD)(CUdeviceptr, const void*, size_t);
  CUresult (*cuMemcpyDtoH)(void*, CUdeviceptr, size_t);
  CUresult (*cuMemcpyHtoDAsync)(CUdeviceptr, const void*, size_t, CUstream);
  CUresult (*cuMemcpyDtoHAsync)(void*, CUdeviceptr, size_t, CUstream);
  CUresult (*cuMemcpy2DAsync)(const CUDA_MEMCPY2D*, CUstream);
  CUresult (*cuMemsetD8Async)(CUdeviceptr, unsigned char, size_t, CUstream);
  CUresult (*cuStreamCreate)(CUstream*, unsigned);
  CUresult (*cuStreamDestroy)(CUstream);
  CUresult (*cuStreamSynchronize)(CUstream);
  CUresult (*cuStreamWaitEvent)(CUstream, CUevent, unsigned);
  CUresult (*cuEventCreate)(CUevent*, unsigned);
  CUresult (*cuEventDestroy)(CUevent);
  CUresult (*cuEventRecord)(CUevent, CUstream);
  CUresult (*cuEventSynchronize)(CUevent);
  CUresult (*cuEventElapsedTime)(float*, CUevent, CUevent);
  CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void**, void**);
  CUresult (*cuFuncGetAttribute)(int*, int, CUfunction);
  CUresult (*cuFuncSetAttribute)(CUfunction, int, int);
  CUresult (*cuGetErrorString)(CUresult, const char**);
};

// returns nullptr (and sets *why) when libcuda cannot be loaded
const Api* api(std::string* why);

}  // namespace cu
}  // namespace rn
