// How fast can host memory be page-locked, and what does it cost the other CUDA calls of the process meanwhile?
//   nvcc -O2 -o /tmp/pin_bench scripts/micro/pin_bench.cu && /tmp/pin_bench
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include <sys/mman.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
	const size_t GB = 1ull << 30;
	cudaFree(0);
	double t = now(); void *p = nullptr; cudaHostAlloc(&p, GB, cudaHostAllocDefault); printf("cudaHostAlloc 1 GB: %.3f s\n", now() - t);
	t = now(); cudaFreeHost(p); printf("cudaFreeHost: %.3f s\n", now() - t);
	// populate in 4 threads, then register
	char *q = (char *)mmap(nullptr, 4 * GB, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	t = now();
	{ std::vector<std::thread> th; for(int i = 0; i < 4; i++) { th.emplace_back([=]() { memset(q + i * GB, 0, GB); }); } for(auto &x : th) { x.join(); } }
	printf("populate 4 GB on 4 threads: %.3f s\n", now() - t);
	t = now(); cudaError_t e = cudaHostRegister(q, GB, cudaHostRegisterDefault); printf("cudaHostRegister 1 GB (populated): %.3f s (%s)\n", now() - t, cudaGetErrorString(e));
	// register 3 more GB on 3 threads at once, while a 4th thread issues small CUDA calls and records their worst latency
	std::atomic<bool> stop(false); double worst = 0; int calls = 0;
	std::thread prober([&]() { void *d = nullptr; while(!stop) { double a = now(); cudaMalloc(&d, 1 << 20); cudaFree(d); double b = now() - a; if(b > worst) { worst = b; } calls++; } });
	t = now();
	{ std::vector<std::thread> th; for(int i = 1; i < 4; i++) { th.emplace_back([=]() { cudaHostRegister(q + i * GB, GB, cudaHostRegisterDefault); }); } for(auto &x : th) { x.join(); } }
	printf("cudaHostRegister 3 x 1 GB on 3 threads: %.3f s; other thread: %d cudaMalloc+cudaFree pairs, worst %.3f s\n", now() - t, calls, worst);
	stop = true; prober.join();
	// copy speed from registered memory
	void *d = nullptr; cudaMalloc(&d, GB); cudaStream_t st; cudaStreamCreate(&st);
	t = now(); cudaMemcpyAsync(d, q, GB, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); printf("H2D 1 GB from registered memory: %.1f GB/s\n", 1.0 / (now() - t));
	stop = false; worst = 0; calls = 0;
	std::thread prober2([&]() { void *d2 = nullptr; while(!stop) { double a = now(); cudaMalloc(&d2, 1 << 20); cudaFree(d2); double b = now() - a; if(b > worst) { worst = b; } calls++; } });
	t = now(); cudaHostAlloc(&p, GB, cudaHostAllocDefault); printf("cudaHostAlloc 1 GB with a prober: %.3f s; prober %d pairs, worst %.3f s\n", now() - t, calls, worst);
	stop = true; prober2.join();
	return 0;
}
