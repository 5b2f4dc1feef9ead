// TEST INFRASTRUCTURE: fiber scheduler behind tests/simt_cpu/shim/cuda_runtime.h.  Blocks run one after another; the threads
// of a block are ucontext fibers resumed round-robin; __syncthreads / warp exchanges park a fiber until its block / warp has
// arrived (threads that returned count as arrived).  A block in which nothing can be released is a deadlock -> abort.
#include "shim/cuda_runtime.h"
#include <ucontext.h>
#include <sys/mman.h>
#include <vector>
#include <stdarg.h>

uint3 threadIdx, blockIdx;
dim3 blockDim(1), gridDim(1);

namespace {
enum State { READY, AT_BLOCK, AT_WARP, DONE };
struct Fiber { ucontext_t ctx; State st; char* stack; };
const size_t kStack = 256 << 10;
std::vector<Fiber> fibers;
std::vector<char*> stacks;
ucontext_t sched_ctx;
int cur = -1, nthreads = 0;
const std::function<void()>* body_fn = nullptr;
std::vector<unsigned long long> slots;                 // one per thread (warp exchange buffers)
std::vector<unsigned char> dyn;                        // dynamic shared memory of the running block
int occupancy = 2;
char last_error[512];

void set_tid(int i) {
  threadIdx.x = i % blockDim.x;
  threadIdx.y = (i / blockDim.x) % blockDim.y;
  threadIdx.z = i / (blockDim.x * blockDim.y);
}
void trampoline() {
  (*body_fn)();
  fibers[cur].st = DONE;
  swapcontext(&fibers[cur].ctx, &sched_ctx);
}
void park(State s) {
  fibers[cur].st = s;
  const int me = cur;
  swapcontext(&fibers[me].ctx, &sched_ctx);
  // resumed by the scheduler with cur == me and threadIdx restored
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
  for (int i = 0; i < n; ++i) {
    getcontext(&fibers[i].ctx);
    fibers[i].ctx.uc_stack.ss_sp = stacks[i];
    fibers[i].ctx.uc_stack.ss_size = kStack;
    fibers[i].ctx.uc_link = &sched_ctx;
    fibers[i].st = READY;
    makecontext(&fibers[i].ctx, trampoline, 0);
  }
  int done = 0;
  while (done < n) {
    bool ran = false;
    for (int i = 0; i < n; ++i) {
      if (fibers[i].st != READY) continue;
      cur = i; set_tid(i);
      swapcontext(&sched_ctx, &fibers[i].ctx);
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

namespace simt {
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body) {
  if (cur != -1) { fprintf(stderr, "simt-cpu: nested launch\n"); abort(); }
  gridDim = grid; blockDim = block;
  nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > 1024) { fprintf(stderr, "simt-cpu: bad block size %d\n", nthreads); abort(); }
  body_fn = &body;
  dyn.assign(dyn_smem_bytes + 16, 0xCD);                                 // poisoned, like uninitialised shared memory
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

// error plumbing of libcolddiff (api.cu is not part of the CPU build)
void cd_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(last_error, sizeof(last_error), fmt, ap); va_end(ap);
}
extern "C" const char* simt_last_error() { return last_error; }
