// TEST INFRASTRUCTURE: fiber scheduler behind tests/simt_cpu/shim/cuda_runtime.h.  Blocks run one after another; the threads
// of a block are ucontext fibers resumed round-robin; __syncthreads / warp exchanges park a fiber until its block / warp has
// arrived (threads that returned count as arrived).  A block in which nothing can be released is a deadlock -> abort.
#include "shim/cuda_runtime.h"
#include <ucontext.h>
#include <sys/mman.h>
#include <vector>
#include <stdarg.h>

extern "C" char __start_simt_shared[] __attribute__((weak));
extern "C" char __stop_simt_shared[] __attribute__((weak));
uint3 threadIdx, blockIdx;
dim3 blockDim(1), gridDim(1);

// ---- context switch: callee-saved registers + stack pointer (no signal-mask system call, unlike swapcontext) ---------------
#if defined(__x86_64__)
extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch, @function
simt_switch:
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
.size simt_switch, .-simt_switch
)");
#define SIMT_FAST_SWITCH 1
#endif

namespace {
enum State { READY, AT_BLOCK, AT_WARP, DONE };
#ifdef SIMT_FAST_SWITCH
struct Fiber { void* sp; State st; };
void* sched_sp;
#else
struct Fiber { ucontext_t ctx; State st; };
ucontext_t sched_ctx;
#endif
const size_t kStack = 256 << 10;
std::vector<Fiber> fibers;
std::vector<char*> stacks;
int cur = -1, nthreads = 0;
const std::function<void()>* body_fn = nullptr;
std::vector<unsigned long long> slots;                 // one per thread (warp exchange buffers)
std::vector<unsigned char> dyn;                        // dynamic shared memory of the running block
int occupancy = 2;
// threads of a block are resumed in ascending order, or descending with SIMT_CPU_ORDER=reverse: a missing barrier between a
// shared-memory write and another thread's read shows as a poisoned (NaN) or stale value under at least one of the two orders
bool reverse_order = getenv("SIMT_CPU_ORDER") && strcmp(getenv("SIMT_CPU_ORDER"), "reverse") == 0;

void set_tid(int i) {
  threadIdx.x = i % blockDim.x;
  threadIdx.y = (i / blockDim.x) % blockDim.y;
  threadIdx.z = i / (blockDim.x * blockDim.y);
}
#ifdef SIMT_FAST_SWITCH
void to_scheduler() { simt_switch(&fibers[cur].sp, sched_sp); }
void to_fiber(int i) { simt_switch(&sched_sp, fibers[i].sp); }
#else
void to_scheduler() { swapcontext(&fibers[cur].ctx, &sched_ctx); }
void to_fiber(int i) { swapcontext(&sched_ctx, &fibers[i].ctx); }
#endif
void trampoline() {
  (*body_fn)();
  fibers[cur].st = DONE;
  to_scheduler();
  abort();                                             // a finished fiber is never resumed
}
void park(State s) {
  fibers[cur].st = s;
  to_scheduler();                                      // resumed by the scheduler with cur and threadIdx restored
}
void run_block() {
  const int n = nthreads;
  if ((int)stacks.size() < n) {
    for (int i = (int)stacks.size(); i < n; ++i) {
      void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (p == MAP_FAILED) { perror("mmap"); abort(); }
      stacks.push_back((char*)p);
    }
  }
  fibers.resize(n);
  slots.assign(n, 0);
  if (__start_simt_shared && __stop_simt_shared > __start_simt_shared)        // static __shared__ arrays: all-ones = NaN
    memset(__start_simt_shared, 0xFF, __stop_simt_shared - __start_simt_shared);
  if (!dyn.empty()) memset(dyn.data(), 0xFF, dyn.size());                      // dynamic shared memory likewise
  for (int i = 0; i < n; ++i) {
    fibers[i].st = READY;
#ifdef SIMT_FAST_SWITCH
    // initial frame: six zeroed callee-saved registers, then the address `ret` jumps to; the slot above it plays the return
    // address of trampoline, so that rsp % 16 == 8 at its entry as the ABI requires
    uintptr_t top = ((uintptr_t)stacks[i] + kStack) & ~(uintptr_t)15;
    void** sp = (void**)(top - 16);
    sp[1] = nullptr;
    sp[0] = (void*)&trampoline;
    for (int r = 1; r <= 6; ++r) sp[-r] = nullptr;
    fibers[i].sp = (void*)(sp - 6);
#else
    getcontext(&fibers[i].ctx);
    fibers[i].ctx.uc_stack.ss_sp = stacks[i];
    fibers[i].ctx.uc_stack.ss_size = kStack;
    fibers[i].ctx.uc_link = &sched_ctx;
    makecontext(&fibers[i].ctx, trampoline, 0);
#endif
  }
  int done = 0;
  while (done < n) {
    bool ran = false;
    for (int k = 0; k < n; ++k) {
      const int i = reverse_order ? n - 1 - k : k;
      if (fibers[i].st != READY) continue;
      cur = i; set_tid(i);
      to_fiber(i);
      ran = true;
      if (fibers[i].st == DONE) ++done;
    }
    bool released = false;
    for (int w = 0; w * 32 < n; ++w) {                                   // warp-level rendezvous
      const int lo = w * 32, hi = std::min(n, lo + 32);
      int waiting = 0, other = 0;
      for (int i = lo; i < hi; ++i) { if (fibers[i].st == AT_WARP) ++waiting; else if (fibers[i].st != DONE) ++other; }
      if (waiting && !other) { for (int i = lo; i < hi; ++i) if (fibers[i].st == AT_WARP) fibers[i].st = READY; released = true; }
    }
    if (!released) {
      int waiting = 0, other = 0;
      for (int i = 0; i < n; ++i) { if (fibers[i].st == AT_BLOCK) ++waiting; else if (fibers[i].st != DONE) ++other; }
      if (waiting && !other) { for (int i = 0; i < n; ++i) if (fibers[i].st == AT_BLOCK) fibers[i].st = READY; released = true; }
    }
    if (!ran && !released && done < n) {
      fprintf(stderr, "simt-cpu: deadlock in block (%u,%u,%u): threads wait at different barriers\n", blockIdx.x, blockIdx.y, blockIdx.z);
      abort();
    }
  }
  cur = -1;
}
}  // namespace

int simt_occupancy() { return occupancy; }
extern "C" void simt_set_occupancy(int n) { occupancy = n; }
extern "C" void simt_set_reverse_order(int r) { reverse_order = r != 0; }

namespace simt {
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body) {
  if (cur != -1) { fprintf(stderr, "simt-cpu: nested launch\n"); abort(); }
  gridDim = grid; blockDim = block;
  nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > 1024) { fprintf(stderr, "simt-cpu: bad block size %d\n", nthreads); abort(); }
  body_fn = &body;
  dyn.assign(dyn_smem_bytes + 16, 0xFF);                                 // NaN-poisoned, like uninitialised shared memory
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
        run_block();
      }
  body_fn = nullptr;
}
void block_barrier() { park(AT_BLOCK); }
void warp_barrier() { park(AT_WARP); }
void* dyn_smem() { return (void*)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15); }
unsigned long long* warp_slot(int lane) { return &slots[(cur / 32) * 32 + lane]; }
int lane_id() { return cur % 32; }
bool lane_alive(int lane) { const int i = (cur / 32) * 32 + lane; return i < nthreads && fibers[i].st != DONE; }
void unsupported(const char* what) { fprintf(stderr, "simt-cpu: unsupported on the CPU: %s\n", what); abort(); }
}  // namespace simt

