// ring.cu -- pinned-host trajectory ring: the learner side of buffer_queue.FIFOQueue
// (distributed_queue/buffer_queue.py:418-512) re-thought for a GPU learner.
//
// The reference dequeues B trajectories one RPC at a time and np.stack()s eight fields into fresh
// pageable arrays (train_impala.py:98-108).  Here the FIFO is organised as batch slots: trajectory i
// of a batch is written at row i of eight field-major [B, ...] arrays inside one pinned block, so a
// popped batch is already "stacked", page-locked and ready for cudaMemcpyAsync -- zero host copies
// between the producer's write and the H2D DMA.  FIFO order is preserved (reservation order).
// Producers (actors) may push concurrently; one consumer (the learner loop) pops.
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace drl;

struct drl_ring {
  int T = 0, H = 0, W = 0, C = 0, A = 0, L = 0, batch = 0, nslots = 0;
  bool pinned = false;
  uint8_t* base = nullptr;
  size_t slot_bytes = 0;
  size_t off_state = 0, off_reward = 0, off_done = 0, off_mu = 0, off_action = 0, off_pa = 0, off_h = 0, off_c = 0;
  size_t sz_state = 0, sz_reward = 0, sz_done = 0, sz_mu = 0, sz_action = 0, sz_pa = 0, sz_h = 0, sz_c = 0;  // per trajectory
  struct SlotState { int reserved = 0, committed = 0; bool held = false; };
  std::vector<SlotState> st;
  int tail = 0, head = 0;
  long long committed_total = 0, popped_total = 0;
  std::mutex mu;
  std::condition_variable cv_space, cv_ready;
};

static size_t ring_align(size_t v) { return (v + 255) / 256 * 256; }

extern "C" {

int drl_ring_create(int32_t trajectory, int32_t height, int32_t width, int32_t channels, int32_t num_action,
                    int32_t lstm_size, int32_t capacity, int32_t batch, int32_t want_pinned, drl_ring** out) {
  if (!out) { set_error("null argument"); return DRL_ERR_INVALID; }
  *out = nullptr;
  if (trajectory < 1 || height < 1 || width < 1 || channels < 1 || num_action < 1 || lstm_size < 1 || batch < 1 ||
      capacity < batch) {
    set_error("ring: invalid dimensions (capacity must be >= batch)");
    return DRL_ERR_INVALID;
  }
  drl_ring* r = new drl_ring();
  r->T = trajectory; r->H = height; r->W = width; r->C = channels; r->A = num_action; r->L = lstm_size;
  r->batch = batch;
  r->nslots = (capacity + batch - 1) / batch + 1;
  const size_t T = trajectory;
  r->sz_state = T * height * width * channels;
  r->sz_reward = T * 4; r->sz_done = T; r->sz_mu = T * num_action * 4; r->sz_action = T * 4; r->sz_pa = T * 4;
  r->sz_h = T * lstm_size * 4; r->sz_c = T * lstm_size * 4;
  size_t off = 0;
  r->off_state = off; off = ring_align(off + batch * r->sz_state);
  r->off_reward = off; off = ring_align(off + batch * r->sz_reward);
  r->off_done = off; off = ring_align(off + batch * r->sz_done);
  r->off_mu = off; off = ring_align(off + batch * r->sz_mu);
  r->off_action = off; off = ring_align(off + batch * r->sz_action);
  r->off_pa = off; off = ring_align(off + batch * r->sz_pa);
  r->off_h = off; off = ring_align(off + batch * r->sz_h);
  r->off_c = off; off = ring_align(off + batch * r->sz_c);
  r->slot_bytes = ring_align(off);
  const size_t total = r->slot_bytes * r->nslots;
  void* p = nullptr;
  if (want_pinned) {
    if (cudaHostAlloc(&p, total, cudaHostAllocPortable) == cudaSuccess) {
      r->pinned = true;
    } else {
      cudaGetLastError();
      p = nullptr;
    }
  }
  if (!p) {
    if (posix_memalign(&p, 4096, total) != 0) {
      delete r;
      set_error("ring: out of host memory (%zu bytes)", total);
      return DRL_ERR_INVALID;
    }
  }
  memset(p, 0, total);
  r->base = static_cast<uint8_t*>(p);
  r->st.resize(r->nslots);
  *out = r;
  return DRL_OK;
}

int drl_ring_destroy(drl_ring* r) {
  if (!r) return DRL_OK;
  if (r->base) {
    if (r->pinned) cudaFreeHost(r->base);
    else free(r->base);
  }
  delete r;
  return DRL_OK;
}

int drl_ring_is_pinned(const drl_ring* r) { return (r && r->pinned) ? 1 : 0; }

int drl_ring_push(drl_ring* r, const uint8_t* state, const float* reward, const uint8_t* done,
                  const float* behavior_policy, const int32_t* action, const int32_t* previous_action,
                  const float* previous_h, const float* previous_c, int32_t timeout_ms) {
  if (!r || !state || !reward || !done || !behavior_policy || !action || !previous_action || !previous_h ||
      !previous_c) {
    set_error("ring push: null argument");
    return DRL_ERR_INVALID;
  }
  int slot, pos;
  {
    std::unique_lock<std::mutex> lk(r->mu);
    auto has_space = [&]() {
      drl_ring::SlotState& s = r->st[r->tail];
      return !s.held && s.reserved < r->batch;
    };
    if (timeout_ms < 0) {
      r->cv_space.wait(lk, has_space);
    } else if (!r->cv_space.wait_for(lk, std::chrono::milliseconds(timeout_ms), has_space)) {
      set_error("ring push: timed out (queue full)");
      return DRL_ERR_TIMEOUT;
    }
    slot = r->tail;
    pos = r->st[slot].reserved++;
    if (r->st[slot].reserved == r->batch) {
      // advance to the next slot; producers block in has_space() while it is still held / unconsumed
      const int next = (r->tail + 1) % r->nslots;
      r->tail = next;
    }
  }
  uint8_t* sb = r->base + (size_t)slot * r->slot_bytes;
  memcpy(sb + r->off_state + pos * r->sz_state, state, r->sz_state);
  memcpy(sb + r->off_reward + pos * r->sz_reward, reward, r->sz_reward);
  memcpy(sb + r->off_done + pos * r->sz_done, done, r->sz_done);
  memcpy(sb + r->off_mu + pos * r->sz_mu, behavior_policy, r->sz_mu);
  memcpy(sb + r->off_action + pos * r->sz_action, action, r->sz_action);
  memcpy(sb + r->off_pa + pos * r->sz_pa, previous_action, r->sz_pa);
  memcpy(sb + r->off_h + pos * r->sz_h, previous_h, r->sz_h);
  memcpy(sb + r->off_c + pos * r->sz_c, previous_c, r->sz_c);
  {
    std::lock_guard<std::mutex> lk(r->mu);
    r->st[slot].committed++;
    r->committed_total++;
    if (r->st[slot].committed == r->batch) r->cv_ready.notify_all();
  }
  return DRL_OK;
}

int drl_ring_pop_batch(drl_ring* r, drl_ring_batch* out, int32_t timeout_ms) {
  if (!r || !out) { set_error("ring pop: null argument"); return DRL_ERR_INVALID; }
  int slot;
  {
    std::unique_lock<std::mutex> lk(r->mu);
    auto ready = [&]() {
      drl_ring::SlotState& s = r->st[r->head];
      return !s.held && s.committed == r->batch;
    };
    if (timeout_ms < 0) {
      r->cv_ready.wait(lk, ready);
    } else if (!r->cv_ready.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready)) {
      set_error("ring pop: timed out (fewer than batch trajectories queued)");
      return DRL_ERR_TIMEOUT;
    }
    slot = r->head;
    r->st[slot].held = true;
    r->head = (r->head + 1) % r->nslots;
    r->popped_total += r->batch;
  }
  uint8_t* sb = r->base + (size_t)slot * r->slot_bytes;
  out->state = sb + r->off_state;
  out->reward = reinterpret_cast<float*>(sb + r->off_reward);
  out->done = sb + r->off_done;
  out->behavior_policy = reinterpret_cast<float*>(sb + r->off_mu);
  out->action = reinterpret_cast<int32_t*>(sb + r->off_action);
  out->previous_action = reinterpret_cast<int32_t*>(sb + r->off_pa);
  out->previous_h = reinterpret_cast<float*>(sb + r->off_h);
  out->previous_c = reinterpret_cast<float*>(sb + r->off_c);
  out->slot = slot;
  return DRL_OK;
}

int drl_ring_release(drl_ring* r, int32_t slot) {
  if (!r || slot < 0 || slot >= r->nslots) { set_error("ring release: bad slot"); return DRL_ERR_INVALID; }
  {
    std::lock_guard<std::mutex> lk(r->mu);
    if (!r->st[slot].held) { set_error("ring release: slot %d is not held", slot); return DRL_ERR_STATE; }
    r->st[slot].held = false;
    r->st[slot].reserved = 0;
    r->st[slot].committed = 0;
  }
  r->cv_space.notify_all();
  return DRL_OK;
}

int drl_ring_size(drl_ring* r) {
  if (!r) return 0;
  std::lock_guard<std::mutex> lk(r->mu);
  return (int)(r->committed_total - r->popped_total);
}

}  // extern "C"
