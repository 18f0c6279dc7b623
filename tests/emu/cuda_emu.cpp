/* cuda_emu.cpp -- fiber scheduler behind cuda_emu.h (test infrastructure, see the header). */
#include "cuda_emu.h"
#include <mutex>

namespace emu {
Fiber *g_cur = nullptr;
Block *g_blk = nullptr;

asm(R"(
.text
.globl emu_swap
.type emu_swap,@function
emu_swap:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size emu_swap,.-emu_swap
)");

static const size_t STACK = 512 * 1024;

/* switch to the next live fiber of the block (round-robin); returns to main when none is left */
static void switch_next(Fiber *from)
{
	Block *b = g_blk;
	unsigned n = (unsigned)b->fibers.size();
	for(unsigned k = 1; k <= n; k++) {
		unsigned i = (b->cur + k) % n;
		if(!b->fibers[i].done) {
			b->cur = i; g_cur = &b->fibers[i];
			if(g_cur != from) { emu_swap(&from->sp, g_cur->sp); }
			return;
		}
	}
	g_cur = nullptr;
	emu_swap(&from->sp, b->main_sp);
}

void yield() { switch_next(g_cur); }

static void complete_warp(Warp *w)
{
	memcpy(w->snap[w->gen & 1], w->slot, sizeof(w->slot));
	w->arrived = 0; w->gen++;
}

const uint64_t *warp_gather(uint64_t v)
{
	Fiber *f = g_cur; Warp *w = f->warp;
	unsigned g = w->gen;
	w->slot[f->lane] = v;
	if(++w->arrived >= w->nlive) { complete_warp(w); }
	else { while(w->gen == g) { yield(); } }
	return w->snap[g & 1];
}

static void trampoline()
{
	Fiber *f = g_cur;
	(*g_blk->body)();
	f->done = 1;
	Warp *w = f->warp;
	w->nlive--; g_blk->nlive--;
	if(w->nlive > 0 && w->arrived >= w->nlive) { complete_warp(w); }
	if(g_blk->nlive > 0 && g_blk->bar_arrived >= g_blk->nlive) { g_blk->bar_arrived = 0; g_blk->bar_gen++; }
	switch_next(f);
	abort();	/* never resumed */
}

void launch(emu_dim3 grid, emu_dim3 block, size_t smem, const std::function<void()> &body)
{
	static std::mutex launch_mu;			/* one kernel at a time: the scheduler state is global (contexts on several host threads share it) */
	std::lock_guard<std::mutex> launch_lock(launch_mu);
	unsigned nt = block.x * block.y * block.z;
	static std::vector<uint8_t *> stacks;
	while(stacks.size() < nt) { stacks.push_back((uint8_t *)aligned_alloc(64, STACK)); }
	Block b;
	b.fibers.resize(nt); b.warps.resize((nt + 31) / 32);
	b.smem = (uint8_t *)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
	b.bdim = block; b.gdim = grid; b.body = &body;
	g_blk = &b;
	for(unsigned bz = 0; bz < grid.z; bz++) for(unsigned by = 0; by < grid.y; by++) for(unsigned bx = 0; bx < grid.x; bx++) {
		b.bid = emu_dim3(bx, by, bz); b.bar_arrived = 0; b.bar_gen = 0; b.nlive = nt; b.cur = nt - 1;
		memset(b.smem, 0xcd, smem);
		for(auto &w : b.warps) { memset(&w, 0, sizeof(Warp)); }
		for(unsigned t = 0; t < nt; t++) {
			Fiber &f = b.fibers[t];
			f.tid = emu_dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
			f.lane = t & 31; f.done = 0; f.warp = &b.warps[t / 32]; f.block = &b; f.stack = stacks[t];
			f.warp->nlive++;
			uint64_t *top = (uint64_t *)(f.stack + STACK - 64);
			top[-2] = (uint64_t)(void *)&trampoline;		/* return address: after `ret`, rsp = top - 8 (== 8 mod 16) */
			f.sp = (void *)(top - 2 - 6);					/* six callee-saved registers below it */
			memset(f.sp, 0, 6 * 8);
		}
		Fiber mainf; memset(&mainf, 0, sizeof(mainf));
		b.cur = 0; g_cur = &b.fibers[0];
		emu_swap(&b.main_sp, g_cur->sp);
	}
	free(b.smem);
	g_blk = nullptr; g_cur = nullptr;
}
}

void __syncthreads()
{
	emu::Block *b = emu::g_blk;
	unsigned g = b->bar_gen;
	if(++b->bar_arrived >= b->nlive) { b->bar_arrived = 0; b->bar_gen++; }
	else { while(b->bar_gen == g) { emu::yield(); } }
}
