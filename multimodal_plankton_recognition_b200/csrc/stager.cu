// Host <-> HBM staging for the loss path (plk_stager_*): a ring of device slots fed by a dedicated
// copy stream, ordered against the consumer's stream with per-slot events.  Plain CUDA runtime calls;
// one C call per step and direction instead of a dozen Python-level stream/event operations.
#include <vector>
#include "common.cuh"

namespace {
struct Stager {
  int device = 0;
  int depth = 0;
  cudaStream_t copy = nullptr;
  std::vector<cudaEvent_t> ready, freed, read;
  std::vector<char> freed_valid, ready_valid;
};
}  // namespace

using namespace plk;

extern "C" {

void* plk_stager_create(int depth, int read_slots) {
  if (depth < 2 || depth > 64 || read_slots < 1 || read_slots > 256) {
    set_error("plk_stager_create: depth must be in [2, 64] and read_slots in [1, 256]");
    return nullptr;
  }
  Stager* s = new Stager();
  s->depth = depth;
  bool ok = cudaGetDevice(&s->device) == cudaSuccess &&
            cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking) == cudaSuccess;
  s->ready.assign(depth, nullptr);
  s->freed.assign(depth, nullptr);
  s->freed_valid.assign(depth, 0);
  s->ready_valid.assign(depth, 0);
  s->read.assign(read_slots, nullptr);
  for (auto* v : {&s->ready, &s->freed, &s->read})
    for (auto& e : *v) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    set_error("plk_stager_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete s;   // events/stream created so far are reclaimed with the context
    return nullptr;
  }
  return s;
}

void plk_stager_destroy(void* h) {
  Stager* s = (Stager*)h;
  if (!s) return;
  for (auto* v : {&s->ready, &s->freed, &s->read})
    for (auto& e : *v)
      if (e) cudaEventDestroy(e);
  if (s->copy) cudaStreamDestroy(s->copy);
  delete s;
}

int plk_stager_issue(void* h, int slot, void* dst_x, const void* src_x, size_t bytes_x, void* dst_y,
                     const void* src_y, size_t bytes_y) {
  Stager* s = (Stager*)h;
  PLK_REQUIRE(s && slot >= 0 && slot < s->depth, PLK_ERR_INVALID, "bad stager handle or slot");
  PLK_REQUIRE(dst_x && src_x && bytes_x > 0, PLK_ERR_INVALID, "null buffer");
  // The slot's PREVIOUS copy read the host buffers the caller is about to let go of (they go back to a
  // pinned-memory pool): wait on the host until that DMA has finished.  It was queued `depth` issues ago,
  // so this returns at once unless the host has run `depth` batches ahead of the copy engine -- in which
  // case it is the back-pressure that keeps a recycled pinned block from being overwritten mid-copy.
  if (s->ready_valid[slot]) PLK_CUDA(cudaEventSynchronize(s->ready[slot]));
  if (s->freed_valid[slot]) PLK_CUDA(cudaStreamWaitEvent(s->copy, s->freed[slot], 0));
  PLK_CUDA(cudaMemcpyAsync(dst_x, src_x, bytes_x, cudaMemcpyHostToDevice, s->copy));
  if (dst_y && bytes_y > 0)
    PLK_CUDA(cudaMemcpyAsync(dst_y, src_y, bytes_y, cudaMemcpyHostToDevice, s->copy));
  PLK_CUDA(cudaEventRecord(s->ready[slot], s->copy));
  s->ready_valid[slot] = 1;
  return PLK_OK;
}

int plk_stager_acquire(void* h, int slot, int release_slot, void* consumer_stream) {
  Stager* s = (Stager*)h;
  PLK_REQUIRE(s && slot >= 0 && slot < s->depth && release_slot < s->depth, PLK_ERR_INVALID,
              "bad stager handle or slot");
  cudaStream_t st = (cudaStream_t)consumer_stream;
  if (release_slot >= 0) {   // everything the consumer queued so far precedes the slot's next rewrite
    PLK_CUDA(cudaEventRecord(s->freed[release_slot], st));
    s->freed_valid[release_slot] = 1;
  }
  PLK_CUDA(cudaStreamWaitEvent(st, s->ready[slot], 0));
  return PLK_OK;
}

int plk_stager_read_async(void* h, int read_slot, void* dst_host, const void* src_dev, size_t bytes,
                          void* consumer_stream) {
  Stager* s = (Stager*)h;
  PLK_REQUIRE(s && read_slot >= 0 && read_slot < (int)s->read.size(), PLK_ERR_INVALID,
              "bad stager handle or read slot");
  PLK_REQUIRE(dst_host && src_dev && bytes > 0, PLK_ERR_INVALID, "null buffer");
  cudaStream_t st = (cudaStream_t)consumer_stream;
  PLK_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
  PLK_CUDA(cudaEventRecord(s->read[read_slot], st));
  return PLK_OK;
}

int plk_stager_read_wait(void* h, int read_slot) {
  Stager* s = (Stager*)h;
  PLK_REQUIRE(s && read_slot >= 0 && read_slot < (int)s->read.size(), PLK_ERR_INVALID,
              "bad stager handle or read slot");
  PLK_CUDA(cudaEventSynchronize(s->read[read_slot]));
  return PLK_OK;
}

}  // extern "C"
