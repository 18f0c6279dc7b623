/*
 * mab_cuda.cu -- the shipped build: CUDA runtime bindings for the host driver (mab_host.inl) and the sm_100a kernels
 * (mab_kernels.cuh).  Compiled in-tree by __graft_entry__.build():
 *   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC -o minialign_b200/libminialign_b200.so
 * There is no CPU fallback: every entry point fails with MAB_ENODEV when the CUDA runtime reports an error.
 */
#include <cuda_runtime.h>
#include "mab_types.h"
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <string>

typedef cudaStream_t RT_STREAM;
typedef cudaEvent_t RT_EVENT;
static thread_local cudaError_t g_rt_last = cudaSuccess;
static inline bool RT_OK(cudaError_t e) { g_rt_last = e; return e == cudaSuccess; }
static inline const char *RT_ERRSTR() { return cudaGetErrorString(g_rt_last); }
static inline cudaError_t RT_SET_DEVICE(int d)				/* mab_init: pick the device and insist on sm_100 */
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if(e != cudaSuccess) { return e; }
	if(d < 0 || d >= n) { return cudaErrorInvalidDevice; }
	e = cudaSetDevice(d);
	if(e != cudaSuccess) { return e; }
	int major = 0;
	e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d);
	if(e != cudaSuccess) { return e; }
	if(major < 10) { return cudaErrorInvalidDevice; }		/* sm_100a code only */
	return cudaSuccess;
}
static inline cudaError_t RT_USE_DEVICE(int d) { return cudaSetDevice(d); }		/* per call: cheap, no property queries (they take driver locks) */
static inline unsigned RT_SM_COUNT(int d) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return (unsigned)n; }
/* persistent extend kernel: 4 CTAs x 4 warps per SM (multiple of the SM count) */
static inline unsigned RT_EXTEND_SLOTS(unsigned n_sm) { return n_sm * 4 * MAB_EXT_CTAS_PER_SM; }
template <class T> static inline cudaError_t RT_MALLOC(T **p, uint64_t n) { return cudaMalloc((void **)p, n); }
template <class T> static inline cudaError_t RT_HOST_ALLOC(T **p, uint64_t n) { return cudaHostAlloc((void **)p, n, cudaHostAllocDefault); }
template <class T> static inline void RT_HOST_FREE(T *p) { if(p) { cudaFreeHost((void *)p); } }
static inline cudaError_t RT_HOST_REGISTER(void *p, uint64_t n) { return cudaHostRegister(p, n, cudaHostRegisterDefault); }
static inline void RT_HOST_UNREGISTER(void *p) { cudaHostUnregister(p); }
template <class T> static inline void RT_FREE(T *p) { if(p) { cudaFree((void *)p); } }
static inline cudaError_t RT_MEMCPY_H2D(void *d, const void *s, uint64_t n) { return cudaMemcpy(d, s, n, cudaMemcpyHostToDevice); }
static inline cudaError_t RT_MEMCPY_D2H(void *d, const void *s, uint64_t n) { return cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost); }
static inline cudaError_t RT_MEMCPY_H2D_ASYNC(void *d, const void *s, uint64_t n, RT_STREAM st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st); }
static inline cudaError_t RT_MEMCPY_D2H_ASYNC(void *d, const void *s, uint64_t n, RT_STREAM st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st); }
static inline cudaError_t RT_MEMSET_ASYNC(void *d, int v, uint64_t n, RT_STREAM st) { return cudaMemsetAsync(d, v, n, st); }
static inline cudaError_t RT_STREAM_CREATE(RT_STREAM *s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); }
static inline cudaError_t RT_STREAM_SYNC(RT_STREAM s) { cudaError_t e = cudaStreamSynchronize(s); if(e == cudaSuccess) { e = cudaGetLastError(); } return e; }
static inline cudaError_t RT_EVENT_CREATE(RT_EVENT *e) { return cudaEventCreate(e); }
static inline void RT_EVENT_DESTROY(RT_EVENT e) { cudaEventDestroy(e); }
static inline void RT_STREAM_DESTROY(RT_STREAM s) { cudaStreamDestroy(s); }
static inline cudaError_t RT_MEM_INFO(size_t *f, size_t *t) { return cudaMemGetInfo(f, t); }
static inline cudaError_t RT_DEVICE_SYNC() { return cudaDeviceSynchronize(); }
static inline cudaError_t RT_EVENT_SYNC(RT_EVENT e) { cudaError_t r = cudaEventSynchronize(e); if(r == cudaSuccess) { r = cudaGetLastError(); } return r; }
static inline cudaError_t RT_SYNC_EVENT_CREATE(RT_EVENT *e) { return cudaEventCreateWithFlags(e, cudaEventBlockingSync | cudaEventDisableTiming); }
static inline void RT_EVENT_RECORD(RT_EVENT e, RT_STREAM s) { cudaEventRecord(e, s); }
static inline void RT_STREAM_WAIT(RT_STREAM s, RT_EVENT e) { cudaStreamWaitEvent(s, e, 0); }
static inline cudaError_t RT_LIGHT_EVENT_CREATE(RT_EVENT *e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); }
static inline float RT_EVENT_MS(RT_EVENT a, RT_EVENT b) { float ms = 0.f; if(cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { ms = 0.f; cudaGetLastError(); } return ms; }
static inline double RT_WALL_MS() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define RT_FUNC_MAX_SMEM(kernel, bytes) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#define RT_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)

#include "mab_host.inl"
