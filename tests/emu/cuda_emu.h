/*
 * cuda_emu.h -- a tiny single-threaded CUDA-on-CPU shim used ONLY by the "not gpu" tests.
 *
 * It lets the product's device source (minialign_b200/csrc/mab_device.cuh) be compiled by g++ and run on the host so
 * the kernels' logic can be checked against the oracle where no GPU exists.  Every CUDA thread of a block is a fiber
 * (own stack, hand-written x86-64 context switch); warp collectives (__shfl_sync, __ballot_sync, ...) are rendezvous
 * points: a lane deposits its value, yields round-robin to its siblings, and resumes when all live lanes of the warp
 * arrived.  Blocks run one after another.  Nothing here is linked into the CUDA build.
 */
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <functional>
#include <vector>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define MAB_EMU 1

struct emu_dim3 { unsigned x, y, z; emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef emu_dim3 dim3;
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };

namespace emu {
struct Warp {
	uint64_t slot[32], snap[2][32];
	unsigned arrived, gen, nlive;
};
struct Block;
struct Fiber {
	void *sp;
	uint8_t *stack;
	emu_dim3 tid;
	unsigned lane, done;
	Warp *warp;
	Block *block;
};
struct Block {
	std::vector<Fiber> fibers;
	std::vector<Warp> warps;
	uint8_t *smem;
	unsigned bar_arrived, bar_gen, nlive;
	emu_dim3 bid, bdim, gdim;
	const std::function<void()> *body;
	void *main_sp;
	unsigned cur;
};
extern Fiber *g_cur;
extern Block *g_blk;
extern "C" void emu_swap(void **save_sp, void *load_sp);
void yield();
void launch(emu_dim3 grid, emu_dim3 block, size_t smem, const std::function<void()> &body);
const uint64_t *warp_gather(uint64_t v);
}

#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_blk->bid)
#define blockDim (emu::g_blk->bdim)
#define gridDim (emu::g_blk->gdim)
#define warpSize 32

/* ---- warp collectives ---- */
static inline unsigned __activemask() { return 0xffffffffu; }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_gather(0); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); const uint64_t *a = emu::warp_gather(x); T r; memcpy(&r, &a[src & 31], sizeof(T)); return r; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); unsigned l = emu::g_cur->lane; const uint64_t *a = emu::warp_gather(x); T r; memcpy(&r, &a[l >= d ? l - d : l], sizeof(T)); return r; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); unsigned l = emu::g_cur->lane; const uint64_t *a = emu::warp_gather(x); T r; memcpy(&r, &a[l + d < 32 ? l + d : l], sizeof(T)); return r; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); unsigned l = emu::g_cur->lane; const uint64_t *a = emu::warp_gather(x); T r; memcpy(&r, &a[(l ^ m) & 31], sizeof(T)); return r; }
static inline unsigned __ballot_sync(unsigned, int p) { const uint64_t *a = emu::warp_gather(p != 0); unsigned r = 0; for(int i = 0; i < 32; i++) { r |= (unsigned)(a[i] & 1) << i; } return r & ((emu::g_cur->warp->nlive >= 32) ? 0xffffffffu : 0xffffffffu); }
static inline unsigned __match_any_sync(unsigned, unsigned v) { const uint64_t *a = emu::warp_gather(v); unsigned r = 0; for(int i = 0; i < 32; i++) { r |= (unsigned)((uint32_t)a[i] == v) << i; } return r; }
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
static inline int __reduce_add_sync(unsigned, int v) { const uint64_t *a = emu::warp_gather((uint64_t)(uint32_t)v); uint32_t s = 0; for(int i = 0; i < 32; i++) { s += (uint32_t)a[i]; } return (int)s; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { const uint64_t *a = emu::warp_gather(v); uint32_t s = 0; for(int i = 0; i < 32; i++) { s += (uint32_t)a[i]; } return s; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { const uint64_t *a = emu::warp_gather(v); uint32_t s = 0; for(int i = 0; i < 32; i++) { s |= (uint32_t)a[i]; } return s; }
static inline int __reduce_max_sync(unsigned, int v) { const uint64_t *a = emu::warp_gather((uint64_t)(uint32_t)v); int s = INT32_MIN; for(int i = 0; i < 32; i++) { s = std::max(s, (int)(uint32_t)a[i]); } return s; }
static inline int __reduce_min_sync(unsigned, int v) { const uint64_t *a = emu::warp_gather((uint64_t)(uint32_t)v); int s = INT32_MAX; for(int i = 0; i < 32; i++) { s = std::min(s, (int)(uint32_t)a[i]); } return s; }
void __syncthreads();

/* ---- scalar intrinsics ---- */
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for(int i = 0; i < 32; i++) { r |= ((x >> i) & 1u) << (31 - i); } return r; }
/* full PTX prmt.b32 (default mode): nibble bit 3 replicates the sign of the selected byte */
static inline unsigned emu_prmt(unsigned a, unsigned b, unsigned s);
/* the CUDA intrinsic masks the selector with 0x7777 (no sign replication) */
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) { return emu_prmt(a, b, s & 0x7777u); }
static inline unsigned emu_prmt(unsigned a, unsigned b, unsigned s)
{
	uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
	for(int i = 0; i < 4; i++) {
		unsigned n = (s >> (4 * i)) & 0xf; unsigned byte = (unsigned)(v >> (8 * (n & 7))) & 0xff;
		if(n & 8) { byte = (byte & 0x80) ? 0xff : 0x00; }
		r |= byte << (8 * i);
	}
	return r;
}
#define EMU_S16X2(name, expr) static inline unsigned name(unsigned a, unsigned b) { unsigned r = 0; for(int i = 0; i < 2; i++) { int x = (int16_t)(a >> (16 * i)), y = (int16_t)(b >> (16 * i)); (void)x; (void)y; r |= ((unsigned)(uint16_t)(expr)) << (16 * i); } return r; }
EMU_S16X2(__vadd2, x + y)
EMU_S16X2(__vsub2, x - y)
EMU_S16X2(__vmaxs2, x > y ? x : y)
EMU_S16X2(__vmins2, x < y ? x : y)
EMU_S16X2(__vcmpeq2, x == y ? 0xffff : 0)
EMU_S16X2(__vminu2, (uint16_t)x < (uint16_t)y ? x : y)
EMU_S16X2(__vmaxu2, (uint16_t)x > (uint16_t)y ? x : y)
EMU_S16X2(__vcmpgts2, x > y ? 0xffff : 0)
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __vimin3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmins2(__vmins2(a, b), c); }
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }
static inline unsigned __viaddmin_s16x2(unsigned a, unsigned b, unsigned c) { return __vmins2(__vadd2(a, b), c); }
static inline unsigned __viaddmin_u16x2(unsigned a, unsigned b, unsigned c) { return __vminu2(__vadd2(a, b), c); }
static inline unsigned __vimin3_u16x2(unsigned a, unsigned b, unsigned c) { return __vminu2(__vminu2(a, b), c); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __ll2float_rn(long long x) { return (float)x; }
static inline float __uint2float_rn(unsigned x) { return (float)x; }
static inline double __uint2double_rn(unsigned x) { return (double)x; }
static inline double __ll2double_rn(long long x) { return (double)x; }
static inline long long __double2ll_rz(double x) { return (x != x) ? (long long)0x8000000000000000ULL : (x >= 9.2233720368547758e18 || x < -9.2233720368547758e18) ? (long long)0x8000000000000000ULL : (long long)x; }
static inline long long __float2ll_rz(float x) { return __double2ll_rz((double)x); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __trap() { fprintf(stderr, "emu: __trap()\n"); abort(); }
static inline unsigned __float2uint_rz(float x) { return (unsigned)x; }
