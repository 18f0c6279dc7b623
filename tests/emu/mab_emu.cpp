/* mab_emu.cpp -- builds the product's device + host source against the CUDA-on-CPU shim (cuda_emu.h) so that the "not gpu"
 * tests can run the very same kernels on the host and compare them with the oracle.  Test infrastructure only; the shipped
 * library is minialign_b200/csrc/mab_cuda.cu. */
#include "cuda_emu.h"
#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int RT_STREAM;
typedef int RT_EVENT;
static inline bool RT_OK(int x) { return x == 0; }
static inline const char *RT_ERRSTR() { return "emu"; }
static inline int RT_SET_DEVICE(int) { return 0; }
static inline int RT_USE_DEVICE(int) { return 0; }
static inline unsigned RT_SM_COUNT(int) { return 2; }
static inline unsigned RT_EXTEND_SLOTS(unsigned n_sm) { return n_sm * 4; }
template <class T> static inline int RT_MALLOC(T **p, uint64_t n) { *p = (T *)aligned_alloc(256, (n + 255) / 256 * 256); memset(*p, 0xab, n); return *p ? 0 : 1; }
template <class T> static inline void RT_FREE(T *p) { free((void *)p); }
template <class T> static inline int RT_HOST_ALLOC(T **p, uint64_t n) { *p = (T *)malloc(n); return *p ? 0 : 1; }
template <class T> static inline void RT_HOST_FREE(T *p) { free((void *)p); }
static inline int RT_HOST_REGISTER(void *, uint64_t) { return 0; }
static inline void RT_HOST_UNREGISTER(void *) {}
static inline int RT_MEMCPY_H2D(void *d, const void *s, uint64_t n) { memcpy(d, s, n); return 0; }
static inline int RT_MEMCPY_D2H(void *d, const void *s, uint64_t n) { memcpy(d, s, n); return 0; }
static inline int RT_MEMCPY_H2D_ASYNC(void *d, const void *s, uint64_t n, RT_STREAM) { memcpy(d, s, n); return 0; }
static inline int RT_MEMCPY_D2H_ASYNC(void *d, const void *s, uint64_t n, RT_STREAM) { memcpy(d, s, n); return 0; }
static inline int RT_MEMSET_ASYNC(void *d, int v, uint64_t n, RT_STREAM) { memset(d, v, n); return 0; }
static inline int RT_STREAM_CREATE(RT_STREAM *s) { *s = 0; return 0; }
static inline int RT_STREAM_SYNC(RT_STREAM) { return 0; }
static inline int RT_EVENT_CREATE(RT_EVENT *e) { *e = 0; return 0; }
static inline void RT_EVENT_DESTROY(RT_EVENT) {}
static inline void RT_STREAM_DESTROY(RT_STREAM) {}
static inline int RT_MEM_INFO(size_t *f, size_t *t) { *f = 64ull << 30; *t = 64ull << 30; return 0; }
static inline int RT_DEVICE_SYNC() { return 0; }
static inline int RT_EVENT_SYNC(RT_EVENT) { return 0; }
static inline int RT_SYNC_EVENT_CREATE(RT_EVENT *e) { *e = 0; return 0; }
static inline void RT_EVENT_RECORD(RT_EVENT, RT_STREAM) {}
static inline void RT_STREAM_WAIT(RT_STREAM, RT_EVENT) {}
static inline int RT_LIGHT_EVENT_CREATE(RT_EVENT *e) { *e = 0; return 0; }
static inline float RT_EVENT_MS(RT_EVENT, RT_EVENT) { return 0.f; }
static inline double RT_WALL_MS() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define RT_FUNC_MAX_SMEM(kernel, bytes) do {} while(0)
#define RT_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(dim3(grid), dim3(block), (smem), [&] { kernel(__VA_ARGS__); })

#include "../../minialign_b200/csrc/mab_host.inl"
