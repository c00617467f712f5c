// Host-buffer engine: what a reference-side `ChunkPreprocessor.run`
// (sai/preprocessors/chunk_preprocessor.py:105-147) calls once per chunk.
//
// The packed tiles are copied host->device in slices on a copy stream while the
// genotype pass (K1) of the previous slice runs on the compute stream; the
// window kernels run once all slices are flagged; results come back in one
// batch of device->host copies.  Device buffers are grow-only and reused across
// calls (one engine per GPU / per worker process).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "host_pack.h"
#include "zt_format.cuh"
#include "zt_simd.h"

struct sai_engine {
  int device = 0;
  cudaStream_t s_copy = nullptr, s_comp = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<cudaEvent_t> ev_dec;    // int8 pipeline, zt wire: slice decoded (its device ring slot is free again)
  int host_threads = 0;               // packer threads of the int8 pipeline (0: hardware concurrency)
  int i8_wire = 0;                    // wire format of the int8 pipeline: 0 = auto, 1 = dense tiles, 2 = zt records
  void* ring = nullptr;               // pinned staging ring of the int8 pipeline
  size_t ring_cap = 0;
  void* h_ztoff = nullptr;            // pinned tile directory of the int8 pipeline's zt records
  size_t h_ztoff_cap = 0;
  uint64_t i8_wire_bytes = 0;         // bytes the last int8 call put on the wire (tiles only)
  // grow-only device buffers
  struct Buf {
    void* p = nullptr;
    size_t cap = 0;
  };
  Buf packed, pos, win, mask, qval, res, cand, counts, sums, hist, neg, dd, zt, ztoff;
  // state of the last call (for sai_engine_rescore_windows)
  int64_t n_sites = 0, W = 0;
  int32_t n_jobs = 0;
  sai_job jobs[SAI_MAX_JOBS];
  sai_layout lay{};  // layout of the resident tiles (sai_engine_score_resident)
};

namespace sai {

static int grow(sai_engine::Buf& b, size_t bytes) {
  if (bytes <= b.cap) return SAI_OK;
  if (b.p) SAI_CUDA_CHECK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    return SAI_E_NOMEM;
  }
  b.cap = want;
  return SAI_OK;
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct ResLayout {
  size_t nsnps, u, q, q_cnt, u_start, q_start, totals, total;
};
static ResLayout res_layout(int64_t W, int n_jobs) {
  ResLayout r{};
  size_t at = 0;
  const size_t n = (size_t)n_jobs * (size_t)W;
  r.nsnps = at, at += align256(sizeof(int32_t) * n);
  r.u = at, at += align256(sizeof(int64_t) * n);
  r.q = at, at += align256(sizeof(double) * n);
  r.q_cnt = at, at += align256(sizeof(int32_t) * n);
  r.u_start = at, at += align256(sizeof(int64_t) * n);
  r.q_start = at, at += align256(sizeof(int64_t) * n);
  r.totals = at, at += align256(sizeof(int64_t) * 2 * n_jobs);
  r.total = at;
  return r;
}

// Window kernel on the device-resident flags + results back to the host.
static int run_windows(sai_engine* e, sai_host_results* out) {
  const int64_t W = e->W;
  const int n_jobs = e->n_jobs;
  const int64_t n_tiles = sai_num_tiles(e->n_sites);
  const int64_t stride = n_tiles * kTile;
  const ResLayout rl = res_layout(W, n_jobs);
  const size_t cu = align256(sizeof(int32_t) * (size_t)n_jobs * out->cap_u);
  const size_t cq = align256(sizeof(int32_t) * (size_t)n_jobs * out->cap_q);
  if (int rc = grow(e->cand, cu + cq + 256)) return rc;
  char* res = static_cast<char*>(e->res.p);
  const uint32_t* d_mask_u = static_cast<const uint32_t*>(e->mask.p);
  const uint32_t* d_mask_q = d_mask_u + (size_t)n_jobs * n_tiles;
  const int64_t* d_ws = static_cast<const int64_t*>(e->win.p);
  int32_t* d_uc = static_cast<int32_t*>(e->cand.p);
  int32_t* d_qc = reinterpret_cast<int32_t*>(static_cast<char*>(e->cand.p) + cu);
  cudaStream_t st = e->s_comp;
  if (int rc = sai_window_stats(
          static_cast<const int32_t*>(e->pos.p), e->n_sites, d_ws, d_ws + W, W, e->jobs, n_jobs,
          d_mask_u, d_mask_q, static_cast<const double*>(e->qval.p), stride,
          reinterpret_cast<int32_t*>(res + rl.nsnps), reinterpret_cast<int64_t*>(res + rl.u),
          reinterpret_cast<double*>(res + rl.q), reinterpret_cast<int32_t*>(res + rl.q_cnt),
          reinterpret_cast<int64_t*>(res + rl.u_start), reinterpret_cast<int64_t*>(res + rl.q_start),
          reinterpret_cast<int64_t*>(res + rl.totals), d_uc, out->cap_u, d_qc, out->cap_q, st))
    return rc;
  const size_t n = (size_t)n_jobs * (size_t)W;
  if (n > 0) {
    SAI_CUDA_CHECK(cudaMemcpyAsync(out->nsnps, res + rl.nsnps, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    SAI_CUDA_CHECK(cudaMemcpyAsync(out->u, res + rl.u, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
    SAI_CUDA_CHECK(cudaMemcpyAsync(out->q, res + rl.q, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    SAI_CUDA_CHECK(cudaMemcpyAsync(out->q_cnt, res + rl.q_cnt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    SAI_CUDA_CHECK(cudaMemcpyAsync(out->u_start, res + rl.u_start, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
    SAI_CUDA_CHECK(cudaMemcpyAsync(out->q_start, res + rl.q_start, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
  }
  SAI_CUDA_CHECK(cudaMemcpyAsync(out->totals, res + rl.totals, sizeof(int64_t) * 2 * n_jobs, cudaMemcpyDeviceToHost, st));
  SAI_CUDA_CHECK(cudaStreamSynchronize(st));
  bool fits = true;
  for (int j = 0; j < n_jobs; ++j) {
    const int64_t tu = out->totals[2 * j], tq = out->totals[2 * j + 1];
    if (tu > out->cap_u || tq > out->cap_q) {
      fits = false;
      continue;
    }
    if (tu > 0)
      SAI_CUDA_CHECK(cudaMemcpyAsync(out->u_cand + (size_t)j * out->cap_u, d_uc + (size_t)j * out->cap_u,
                                     sizeof(int32_t) * tu, cudaMemcpyDeviceToHost, st));
    if (tq > 0)
      SAI_CUDA_CHECK(cudaMemcpyAsync(out->q_cand + (size_t)j * out->cap_q, d_qc + (size_t)j * out->cap_q,
                                     sizeof(int32_t) * tq, cudaMemcpyDeviceToHost, st));
  }
  SAI_CUDA_CHECK(cudaStreamSynchronize(st));
  if (!fits) {
    set_error("candidate capacity too small (see totals)");
    return SAI_E_CAPACITY;
  }
  return SAI_OK;
}

static int check_results(const sai_host_results* out) {
  SAI_REQUIRE(out && out->nsnps && out->u && out->q && out->q_cnt && out->u_start && out->q_start &&
                  out->totals,
              "NULL result pointer");
  SAI_REQUIRE(out->cap_u >= 0 && out->cap_q >= 0 && (out->cap_u == 0 || out->u_cand) &&
                  (out->cap_q == 0 || out->q_cand),
              "bad candidate buffers");
  return SAI_OK;
}

}  // namespace sai

using namespace sai;

extern "C" {

int sai_engine_create(int32_t device, sai_engine** out) {
  SAI_REQUIRE(out, "NULL argument");
  SAI_CUDA_CHECK(cudaSetDevice(device));
  sai_engine* e = new sai_engine();
  e->device = device;
  SAI_CUDA_CHECK(cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
  SAI_CUDA_CHECK(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking));
  *out = e;
  return SAI_OK;
}

void sai_engine_destroy(sai_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (auto* b : {&e->packed, &e->pos, &e->win, &e->mask, &e->qval, &e->res, &e->cand, &e->counts, &e->sums, &e->hist, &e->neg, &e->dd, &e->zt, &e->ztoff})
    if (b->p) cudaFree(b->p);
  for (auto ev : e->ev) cudaEventDestroy(ev);
  for (auto ev : e->ev_dec) cudaEventDestroy(ev);
  if (e->ring) cudaFreeHost(e->ring);
  if (e->h_ztoff) cudaFreeHost(e->h_ztoff);
  if (e->s_copy) cudaStreamDestroy(e->s_copy);
  if (e->s_comp) cudaStreamDestroy(e->s_comp);
  delete e;
}

// int8 pipeline (sai_engine_score_host_i8): the reference-side representation -- one int8 matrix
// of per-individual allele sums per population, pageable host memory -- is packed into tiled
// bit-planes by a pool of host threads, slice by slice, straight into a ring of pinned staging
// buffers; as soon as a slice is complete it goes H2D and, once landed, through the genotype
// pass, while the threads are already packing the following slices.  Three stages overlap: pack
// (host cores, reads 4x the bytes it writes) | PCIe copy | K1.
//
// Wire format (sai_engine_set_i8_wire).  The packers saturate the host's memory system, so every
// byte they do not write -- and the copy engine does not read back -- is throughput.  By default a
// packer keeps the tile it has just packed in its L1, turns it into a zt record on the spot
// (zt_simd.cpp) and streams the records of its block of tiles, back to back, to the block's fixed
// region of the ring slot (region = where the dense tiles would have gone, so no packer waits for
// another one's length).  A slice then goes out as ONE strided copy (rows = blocks, width = the
// longest block) into a device ring laid out like the pinned one, followed by the slice's part of
// the tile directory; k_zt_decode rebuilds the dense tiles in HBM in front of K1 and frees the
// device slot.
struct I8Source {
  const int8_t* const* gt;
  const int64_t* row_stride;
};

static int stream_tiles_from_i8(sai_engine* e, const sai_layout* lay, const I8Source& src, int64_t n_sites,
                                int64_t n_tiles, const sai_job* jobs, int32_t n_jobs, uint32_t* d_mask_u,
                                uint32_t* d_mask_q, double* d_qval, int64_t stride) {
  const size_t tile_bytes = (size_t)lay->pairs_per_site * kTile * 8;
  uint64_t slice_bytes = 32ull << 20;
  int kRing = 4;
#ifdef SAI_EXPERIMENTS
  if (const char* v = getenv("SAI_I8_SLICE_MB")) slice_bytes = (uint64_t)std::max(1, atoi(v)) << 20;  // tools/ A/B knobs
  if (const char* v = getenv("SAI_I8_RING")) kRing = std::max(2, atoi(v));
#endif
  // a small chunk does not need (or pay for) the full ring: at most n_tiles / kRing tiles per slot
  const int64_t slice_tiles = std::max<int64_t>(1, std::min<int64_t>((int64_t)(slice_bytes / tile_bytes), (n_tiles + kRing - 1) / kRing));
  const int64_t n_slices = (n_tiles + slice_tiles - 1) / slice_tiles;
  const size_t slot_bytes = (size_t)slice_tiles * tile_bytes;
  if (e->ring_cap < slot_bytes * kRing) {
    if (e->ring) SAI_CUDA_CHECK(cudaFreeHost(e->ring));
    e->ring = nullptr;
    e->ring_cap = 0;
    SAI_CUDA_CHECK(cudaHostAlloc(&e->ring, slot_bytes * kRing, cudaHostAllocDefault));
    e->ring_cap = slot_bytes * kRing;
  }
  while ((int64_t)e->ev.size() < kRing) {
    cudaEvent_t ev;
    SAI_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->ev.push_back(ev);
  }
  // auto: zt records when this CPU has the vector record encoder (the portable one would be the bottleneck)
  const bool zt = e->i8_wire == 2 || (e->i8_wire == 0 && strcmp(zt_isa(), "portable") != 0);
  const int P = lay->pairs_per_site;
  std::vector<uint64_t> padc(P);
  uint64_t* h_off = nullptr;
  if (zt) {
    for (int r = 0; r < P; ++r) padc[r] = pad_constant(*lay, r);
    const size_t need = sizeof(uint64_t) * (size_t)(n_tiles + 1);
    if (e->h_ztoff_cap < need) {
      if (e->h_ztoff) SAI_CUDA_CHECK(cudaFreeHost(e->h_ztoff));
      e->h_ztoff = nullptr;
      e->h_ztoff_cap = 0;
      SAI_CUDA_CHECK(cudaHostAlloc(&e->h_ztoff, need + need / 4, cudaHostAllocDefault));
      e->h_ztoff_cap = need + need / 4;
    }
    h_off = static_cast<uint64_t*>(e->h_ztoff);
    // the records live in a device ring laid out like the pinned one: a slot is rewritten once its
    // slice has been decoded (ev_dec)
    if (int k = grow(e->zt, slot_bytes * kRing + 256)) return k;
    if (int k = grow(e->ztoff, need + 256)) return k;
    while ((int64_t)e->ev_dec.size() < kRing) {
      cudaEvent_t ev;
      SAI_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->ev_dec.push_back(ev);
    }
  }
  int n_threads = e->host_threads > 0 ? e->host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  // packing tasks: blocks of tiles, in slice order
  int64_t block_cap = 32;
#ifdef SAI_EXPERIMENTS
  if (const char* v = getenv("SAI_I8_BLOCK_TILES")) block_cap = std::max(1, atoi(v));
#endif
  const int64_t block_tiles = std::max<int64_t>(1, std::min<int64_t>(block_cap, (slice_tiles + 2 * n_threads - 1) / (2 * n_threads)));
  const int64_t blocks_per_slice = (slice_tiles + block_tiles - 1) / block_tiles;
  auto blocks_of = [&](int64_t s) {
    const int64_t tiles = std::min(slice_tiles, n_tiles - s * slice_tiles);
    return (tiles + block_tiles - 1) / block_tiles;
  };
  const int64_t n_tasks = (n_slices - 1) * blocks_per_slice + blocks_of(n_slices - 1);
  n_threads = (int)std::min<int64_t>(n_threads, n_tasks);
  std::unique_ptr<uint32_t[]> used(new uint32_t[n_tasks]);  // zt: bytes of block i's records (whole 64-byte lines)
  std::atomic<uint64_t> wire_bytes{0};
  std::atomic<int64_t> next_task{0};
  std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[n_slices]);
  for (int64_t s = 0; s < n_slices; ++s) done[s].store(0);
  std::atomic<int> bad{0}, abort_flag{0};
  uint8_t* ring = static_cast<uint8_t*>(e->ring);
  // Slices are handed to the GPU by whichever packer thread completes them (in slice order, under
  // issue_mu): no hand-off to a feeding thread, so no wake-up latency between "packed" and "on the
  // wire", and nothing but copies in the copy stream (a host callback there would stall the next
  // copy until a CPU is free to run it -- all of them are packing).  A ring slot is free again when
  // the copy event of the slice that used it last has completed; the packers check that themselves
  // (`slot_free`), which in the pack-bound steady state is a single relaxed load.
  std::mutex issue_mu;
  int64_t next_issue = 0;          // guarded by issue_mu
  std::atomic<int64_t> issued{0};  // == next_issue, readable without the lock
  std::atomic<int64_t> freed{0};   // copies [0, freed) have left their ring slots
  std::mutex free_mu;
  int rc = SAI_OK;      // guarded by issue_mu
  std::string err_msg;  // the failing thread's message (sai_last_error is thread-local)
  auto fail_locked = [&](int code) {
    if (rc == SAI_OK) {
      rc = code;
      err_msg = sai_last_error();
    }
    abort_flag.store(1);
  };
  auto slot_free = [&](int64_t s) {  // may slice s be packed into its ring slot?
    if (s < kRing || freed.load(std::memory_order_acquire) > s - kRing) return true;
    std::lock_guard<std::mutex> lk(free_mu);
    int64_t f = freed.load(std::memory_order_relaxed);
    // the event of slot f % kRing still belongs to slice f: slice f + kRing is not packed before f is freed
    while (f < issued.load(std::memory_order_acquire) && cudaEventQuery(e->ev[f % kRing]) == cudaSuccess) ++f;
    freed.store(f, std::memory_order_release);
    return f > s - kRing;
  };
  auto issue_ready_slices = [&]() {
    std::lock_guard<std::mutex> lk(issue_mu);
    while (rc == SAI_OK && next_issue < n_slices &&
           done[next_issue].load(std::memory_order_acquire) == (int)blocks_of(next_issue)) {
      const int64_t s = next_issue;
      if (bad.load()) {
        fail_locked(SAI_E_DOMAIN);
        break;
      }
      const int64_t t0 = s * slice_tiles, t1 = std::min(n_tiles, t0 + slice_tiles);
      const int slot = (int)(s % kRing);
      cudaError_t ce = cudaSuccess;
      if (!zt) {
        ce = cudaMemcpyAsync(static_cast<char*>(e->packed.p) + (size_t)t0 * tile_bytes, ring + (size_t)slot * slot_bytes,
                             (size_t)(t1 - t0) * tile_bytes, cudaMemcpyHostToDevice, e->s_copy);
        wire_bytes.fetch_add((uint64_t)(t1 - t0) * tile_bytes, std::memory_order_relaxed);
      } else {
        // rows = the slice's blocks (pitch = a block's dense size), width = the longest record run;
        // a shorter last block goes separately so that no row reaches past the slot
        const int64_t nb = blocks_of(s), task0 = s * blocks_per_slice;
        const bool short_last = (t1 - t0) % block_tiles != 0;
        const int64_t rows = short_last ? nb - 1 : nb;
        const size_t pitch = (size_t)block_tiles * tile_bytes;
        size_t width = 0;
        for (int64_t b = 0; b < rows; ++b) width = std::max<size_t>(width, used[task0 + b]);
        char* dst = static_cast<char*>(e->zt.p) + (size_t)slot * slot_bytes;
        const uint8_t* srcp = ring + (size_t)slot * slot_bytes;
        if (s >= kRing) ce = cudaStreamWaitEvent(e->s_copy, e->ev_dec[slot], 0);  // the slot's previous slice is decoded
        if (ce == cudaSuccess && rows > 0 && width > 0)
          ce = cudaMemcpy2DAsync(dst, pitch, srcp, pitch, width, (size_t)rows, cudaMemcpyHostToDevice, e->s_copy);
        if (ce == cudaSuccess && short_last && used[task0 + nb - 1] > 0)
          ce = cudaMemcpyAsync(dst + (size_t)rows * pitch, srcp + (size_t)rows * pitch, used[task0 + nb - 1],
                               cudaMemcpyHostToDevice, e->s_copy);
        if (ce == cudaSuccess)
          ce = cudaMemcpyAsync(static_cast<uint64_t*>(e->ztoff.p) + t0, h_off + t0, sizeof(uint64_t) * (size_t)(t1 - t0),
                               cudaMemcpyHostToDevice, e->s_copy);
        wire_bytes.fetch_add((uint64_t)width * rows + (short_last ? used[task0 + nb - 1] : 0) + 8ull * (t1 - t0),
                             std::memory_order_relaxed);
      }
      if (ce == cudaSuccess) ce = cudaEventRecord(e->ev[slot], e->s_copy);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->s_comp, e->ev[slot], 0);
      if (ce != cudaSuccess) {
        set_error("int8 pipeline: CUDA call failed: %s", cudaGetErrorString(ce));
        fail_locked(SAI_E_CUDA);
        break;
      }
      ++next_issue;
      issued.store(next_issue, std::memory_order_release);
      if (zt) {
        int k = sai_zt_decode(lay, e->zt.p, (uint64_t)slot_bytes * kRing, static_cast<const uint64_t*>(e->ztoff.p), t0, t1 - t0,
                              e->packed.p, e->s_comp);
        if (k == SAI_OK && cudaEventRecord(e->ev_dec[slot], e->s_comp) != cudaSuccess) {
          set_error("int8 pipeline: cudaEventRecord failed");
          k = SAI_E_CUDA;
        }
        if (k) {
          fail_locked(k);
          break;
        }
      }
      if (int k = sai_site_flags(lay, e->packed.p, t0, t1 - t0, n_tiles, jobs, n_jobs, d_mask_u, d_mask_q, d_qval,
                                 stride, nullptr, nullptr, 0, 0, e->s_comp))
        fail_locked(k);
    }
  };
  const int device = e->device;
  auto worker = [&]() {
    cudaSetDevice(device);  // a fresh thread starts on device 0
    std::unique_ptr<ZtBlockScratch> scratch(zt ? new ZtBlockScratch(P) : nullptr);
    for (;;) {
      const int64_t i = next_task.fetch_add(1);
      if (i >= n_tasks) break;
      const int64_t s = i / blocks_per_slice, b = i % blocks_per_slice;
      while (!slot_free(s) && !abort_flag.load())  // the slot's previous slice is still on the wire (rare:
        std::this_thread::sleep_for(std::chrono::microseconds(50));  // the copy is faster than the packers)
      if (abort_flag.load()) break;
      const int64_t t0 = s * slice_tiles + b * block_tiles;
      const int64_t t1 = std::min(std::min(n_tiles, (s + 1) * slice_tiles), t0 + block_tiles);
      uint8_t* slot = ring + (size_t)(s % kRing) * slot_bytes;
      if (zt) {
        // the block's records -> its region of the ring slot (= where its dense tiles would go); the device
        // ring is laid out like the pinned one, so a record's device offset is its offset in the ring
        uint8_t* region = slot + (size_t)(t0 - s * slice_tiles) * tile_bytes;
        bool bad_here = false;
        used[i] = (uint32_t)zt_pack_block_i8(*lay, src.gt, src.row_stride, n_sites, t0, t1, padc.data(), region,
                                             (uint64_t)(region - ring), h_off, *scratch, true, &bad_here);
        if (bad_here) bad.store(1);
      } else if (pack_tiles_i8_all(*lay, src.gt, src.row_stride, n_sites, t0, t1, s * slice_tiles, slot, 0)) {
        bad.store(1);
      }
      if (done[s].fetch_add(1, std::memory_order_acq_rel) + 1 == (int)blocks_of(s)) issue_ready_slices();
    }
  };
  std::vector<std::thread> pool;
  for (int i = 0; i < n_threads; ++i) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  issue_ready_slices();  // nothing left unless a packer bailed out
  const cudaError_t drained = cudaStreamSynchronize(e->s_copy);  // the ring may be reused by the next call
  e->i8_wire_bytes = wire_bytes.load();
  if (rc == SAI_E_DOMAIN || (rc == SAI_OK && bad.load())) {
    set_error("a genotype value does not fit the bit-planes of its population");
    return SAI_E_DOMAIN;
  }
  if (rc == SAI_OK && drained != cudaSuccess) {
    set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(drained));
    return SAI_E_CUDA;
  }
  if (rc != SAI_OK) set_error("%s", err_msg.c_str());
  return rc;
}

// Shared body of the host entry points: `packed` (dense tiles), the zero-suppressed stream +
// its tile directory when `zt_stream` is given, or int8 matrices (`i8`).
static int score_host_impl(sai_engine* e, const sai_layout* lay, const uint8_t* packed,
                           const uint8_t* zt_stream, const uint64_t* zt_off, const I8Source* i8, const int32_t* pos,
                           int64_t n_sites, const int64_t* win_start, const int64_t* win_end,
                           int64_t n_windows, const sai_job* jobs, int32_t n_jobs, sai_host_results* out) {
  SAI_REQUIRE(e, "NULL engine");
  if (int rc = validate_layout(lay)) return rc;
  if (int rc = validate_jobs(lay, jobs, n_jobs)) return rc;
  SAI_REQUIRE(n_sites >= 0 && n_windows >= 0, "negative size");
  SAI_REQUIRE(n_sites == 0 || ((packed || (zt_stream && zt_off) || i8) && pos), "NULL input");
  SAI_REQUIRE(n_windows == 0 || (win_start && win_end), "NULL windows");
  if (int rc = check_results(out)) return rc;
  SAI_CUDA_CHECK(cudaSetDevice(e->device));
  const int64_t W = n_windows;
  const int64_t n_tiles = sai_num_tiles(n_sites);
  const int64_t stride = n_tiles * kTile;
  const size_t packed_bytes = sai_packed_bytes(lay, n_sites);
  const ResLayout rl = res_layout(W, n_jobs);
  const bool zt = zt_stream != nullptr && n_sites > 0;
  const uint64_t kRaw = 1ull << 63;
  const uint64_t zt_bytes = zt ? (zt_off[n_tiles] & ~kRaw) : 0;

  if (int rc = grow(e->packed, packed_bytes + 256)) return rc;
  if (int rc = grow(e->pos, sizeof(int32_t) * (size_t)stride + 256)) return rc;
  if (int rc = grow(e->win, sizeof(int64_t) * 2 * (size_t)W + 256)) return rc;
  if (int rc = grow(e->mask, sizeof(uint32_t) * 2 * (size_t)n_jobs * n_tiles + 256)) return rc;
  if (int rc = grow(e->qval, sizeof(double) * (size_t)n_jobs * stride + 256)) return rc;
  if (int rc = grow(e->res, rl.total + 256)) return rc;
  if (zt) {
    if (int rc = grow(e->zt, zt_bytes + 256)) return rc;
    if (int rc = grow(e->ztoff, sizeof(uint64_t) * (size_t)(n_tiles + 1) + 256)) return rc;
  }
  e->n_sites = n_sites;
  e->W = W;
  e->n_jobs = n_jobs;
  e->lay = *lay;
  for (int j = 0; j < n_jobs; ++j) e->jobs[j] = jobs[j];

  uint32_t* d_mask_u = static_cast<uint32_t*>(e->mask.p);
  uint32_t* d_mask_q = d_mask_u + (size_t)n_jobs * n_tiles;
  double* d_qval = static_cast<double*>(e->qval.p);
  int32_t* d_pos = static_cast<int32_t*>(e->pos.p);
  int64_t* d_ws = static_cast<int64_t*>(e->win.p);

  // small inputs first
  if (n_sites > 0)
    SAI_CUDA_CHECK(cudaMemcpyAsync(d_pos, pos, sizeof(int32_t) * n_sites, cudaMemcpyHostToDevice, e->s_copy));
  if (W > 0) {
    SAI_CUDA_CHECK(cudaMemcpyAsync(d_ws, win_start, sizeof(int64_t) * W, cudaMemcpyHostToDevice, e->s_copy));
    SAI_CUDA_CHECK(cudaMemcpyAsync(d_ws + W, win_end, sizeof(int64_t) * W, cudaMemcpyHostToDevice, e->s_copy));
  }
  if (zt)
    SAI_CUDA_CHECK(cudaMemcpyAsync(e->ztoff.p, zt_off, sizeof(uint64_t) * (size_t)(n_tiles + 1),
                                   cudaMemcpyHostToDevice, e->s_copy));
  if (i8 && n_tiles > 0) {
    if (int rc = stream_tiles_from_i8(e, lay, *i8, n_sites, n_tiles, jobs, n_jobs, d_mask_u, d_mask_q, d_qval, stride))
      return rc;
    while (e->ev.empty()) {
      cudaEvent_t ev;
      SAI_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->ev.push_back(ev);
    }
    SAI_CUDA_CHECK(cudaEventRecord(e->ev[0], e->s_copy));  // positions / windows went on the copy stream
    SAI_CUDA_CHECK(cudaStreamWaitEvent(e->s_comp, e->ev[0], 0));
    return run_windows(e, out);
  }
  // sliced H2D of the tiles (about 32 MB on the wire per slice); per slice, as soon as it has
  // landed: [zt: rebuild the dense tiles,] then K1
  const size_t tile_bytes = (size_t)lay->pairs_per_site * kTile * 8;
  const uint64_t slice_bytes = 32ull << 20;
  std::vector<int64_t> cut{0};
  if (!zt) {
    const int64_t slice_tiles = std::max<int64_t>(1, (int64_t)(slice_bytes / tile_bytes));
    for (int64_t t = slice_tiles; t < n_tiles; t += slice_tiles) cut.push_back(t);
  } else {
    uint64_t start = 0;
    for (int64_t t = 1; t < n_tiles; ++t) {
      const uint64_t o = zt_off[t] & ~kRaw;
      SAI_REQUIRE(o >= (zt_off[t - 1] & ~kRaw) && o <= zt_bytes, "zt tile directory is not monotonic");
      if (o - start >= slice_bytes) {
        cut.push_back(t);
        start = o;
      }
    }
  }
  if (n_tiles > 0) cut.push_back(n_tiles);
  const int64_t n_slices = (int64_t)cut.size() - 1;
  while ((int64_t)e->ev.size() < n_slices + 1) {
    cudaEvent_t ev;
    SAI_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->ev.push_back(ev);
  }
  for (int64_t s = 0; s < n_slices; ++s) {
    const int64_t t0 = cut[s], t1 = cut[s + 1];
    if (!zt) {
      SAI_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(e->packed.p) + t0 * tile_bytes,
                                     packed + t0 * tile_bytes, (size_t)(t1 - t0) * tile_bytes,
                                     cudaMemcpyHostToDevice, e->s_copy));
    } else {
      const uint64_t b0 = zt_off[t0] & ~kRaw, b1 = zt_off[t1] & ~kRaw;
      if (b1 > b0)
        SAI_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(e->zt.p) + b0, zt_stream + b0, b1 - b0,
                                       cudaMemcpyHostToDevice, e->s_copy));
    }
    SAI_CUDA_CHECK(cudaEventRecord(e->ev[s], e->s_copy));
    SAI_CUDA_CHECK(cudaStreamWaitEvent(e->s_comp, e->ev[s], 0));
    if (zt)
      if (int rc = sai_zt_decode(lay, e->zt.p, zt_bytes, static_cast<const uint64_t*>(e->ztoff.p), t0, t1 - t0,
                                 e->packed.p, e->s_comp))
        return rc;
    if (int rc = sai_site_flags(lay, e->packed.p, t0, t1 - t0, n_tiles, jobs, n_jobs, d_mask_u,
                                d_mask_q, d_qval, stride, nullptr, nullptr, 0, 0, e->s_comp))
      return rc;
  }
  SAI_CUDA_CHECK(cudaEventRecord(e->ev[n_slices], e->s_copy));
  SAI_CUDA_CHECK(cudaStreamWaitEvent(e->s_comp, e->ev[n_slices], 0));
  return run_windows(e, out);
}

int sai_engine_score_host(sai_engine* e, const sai_layout* lay, const uint8_t* packed,
                          const int32_t* pos, int64_t n_sites, const int64_t* win_start,
                          const int64_t* win_end, int64_t n_windows, const sai_job* jobs,
                          int32_t n_jobs, sai_host_results* out) {
  SAI_REQUIRE(n_sites <= 0 || packed, "NULL input");
  return score_host_impl(e, lay, packed, nullptr, nullptr, nullptr, pos, n_sites, win_start, win_end, n_windows, jobs,
                         n_jobs, out);
}

int sai_engine_score_host_zt(sai_engine* e, const sai_layout* lay, const uint8_t* zt_stream,
                             const uint64_t* zt_tile_off, const int32_t* pos, int64_t n_sites,
                             const int64_t* win_start, const int64_t* win_end, int64_t n_windows,
                             const sai_job* jobs, int32_t n_jobs, sai_host_results* out) {
  SAI_REQUIRE(n_sites <= 0 || (zt_stream && zt_tile_off), "NULL input");
  return score_host_impl(e, lay, nullptr, zt_stream, zt_tile_off, nullptr, pos, n_sites, win_start, win_end, n_windows,
                         jobs, n_jobs, out);
}

int sai_engine_score_host_i8(sai_engine* e, const sai_layout* lay, const int8_t* const* gt,
                             const int64_t* row_stride, const int32_t* pos, int64_t n_sites,
                             const int64_t* win_start, const int64_t* win_end, int64_t n_windows,
                             const sai_job* jobs, int32_t n_jobs, sai_host_results* out) {
  SAI_REQUIRE(lay && gt && row_stride, "NULL input");
  if (int rc = validate_layout(lay)) return rc;
  for (int p = 0; p < lay->n_pops; ++p) {
    SAI_REQUIRE(n_sites <= 0 || gt[p], "NULL genotype matrix of population %d", p);
    SAI_REQUIRE(row_stride[p] >= lay->pop[p].n_samples, "row_stride of population %d smaller than n_samples", p);
  }
  const I8Source src{gt, row_stride};
  return score_host_impl(e, lay, nullptr, nullptr, nullptr, &src, pos, n_sites, win_start, win_end, n_windows, jobs,
                         n_jobs, out);
}

int sai_engine_set_i8_wire(sai_engine* e, int32_t mode) {
  SAI_REQUIRE(e, "NULL engine");
  SAI_REQUIRE(mode >= 0 && mode <= 2, "wire format: 0 = auto, 1 = dense tiles, 2 = zt records");
  e->i8_wire = mode;
  return SAI_OK;
}

uint64_t sai_engine_i8_wire_bytes(const sai_engine* e) { return e ? e->i8_wire_bytes : 0; }

int sai_engine_set_host_threads(sai_engine* e, int32_t n_threads) {
  SAI_REQUIRE(e, "NULL engine");
  e->host_threads = n_threads > 0 ? n_threads : 0;
  return SAI_OK;
}

int sai_engine_pattern_sums(sai_engine* e, const sai_layout* lay, int32_t ref_pop, int32_t tgt_pop,
                            int32_t out_pop, const int32_t* src_pops, int32_t n_src, double* sums) {
  SAI_REQUIRE(e && e->n_jobs > 0, "no previous sai_engine_score_host call");
  SAI_REQUIRE(sums && src_pops, "NULL argument");
  if (int rc = validate_layout(lay)) return rc;
  SAI_CUDA_CHECK(cudaSetDevice(e->device));
  const int64_t n_tiles = sai_num_tiles(e->n_sites);
  const int64_t stride = n_tiles * kTile;
  const int64_t W = e->W;
  if (W == 0 || n_src <= 0) return SAI_OK;
  if (int rc = grow(e->counts, sizeof(int32_t) * 2 * (size_t)lay->n_pops * stride + 256)) return rc;
  if (int rc = grow(e->sums, sizeof(double) * 7 * (size_t)n_src * W + 256)) return rc;
  int32_t* d_num = static_cast<int32_t*>(e->counts.p);
  int32_t* d_called = d_num + (size_t)lay->n_pops * stride;
  if (int rc = sai_site_counts(lay, e->packed.p, 0, n_tiles, d_num, d_called, stride, 0, e->s_comp)) return rc;
  const int64_t* d_ws = static_cast<const int64_t*>(e->win.p);
  if (int rc = sai_window_patterns(lay, static_cast<const int32_t*>(e->pos.p), e->n_sites, d_ws, d_ws + W, W,
                                   d_num, d_called, stride, ref_pop, tgt_pop, out_pop, src_pops, n_src,
                                   static_cast<double*>(e->sums.p), e->s_comp))
    return rc;
  SAI_CUDA_CHECK(cudaMemcpyAsync(sums, e->sums.p, sizeof(double) * 7 * (size_t)n_src * W,
                                 cudaMemcpyDeviceToHost, e->s_comp));
  SAI_CUDA_CHECK(cudaStreamSynchronize(e->s_comp));
  return SAI_OK;
}

int sai_engine_dd_sums(sai_engine* e, const sai_layout* lay, int32_t ref_pop, int32_t tgt_pop,
                       const int32_t* src_pops, int32_t n_src, const int64_t* neg_off,
                       const int32_t* neg_site, const int32_t* neg_ind, const int32_t* neg_val,
                       int64_t* ref_sum, int64_t* tgt_sum, int32_t m_max) {
  SAI_REQUIRE(e && e->n_jobs > 0, "no previous sai_engine_score_host call");
  SAI_REQUIRE(ref_sum && tgt_sum && src_pops && neg_off, "NULL argument");
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(ref_pop >= 0 && ref_pop < lay->n_pops && tgt_pop >= 0 && tgt_pop < lay->n_pops,
              "bad population index");
  SAI_REQUIRE(m_max >= 1, "m_max must be positive");
  SAI_CUDA_CHECK(cudaSetDevice(e->device));
  const int64_t n_tiles = sai_num_tiles(e->n_sites);
  const int64_t stride = n_tiles * kTile;
  const int64_t W = e->W;
  if (W == 0 || n_src <= 0) return SAI_OK;
  const int64_t n_neg = neg_off[lay->n_pops];
  SAI_REQUIRE(n_neg >= 0 && (n_neg == 0 || (neg_site && neg_ind && neg_val)), "bad negative-value table");
  // histograms of ref and tgt (what the distance kernels read) and, behind them, of the source
  // populations: only their missing-call totals are used, to check the table against the tiles
  SAI_REQUIRE(n_src <= SAI_MAX_SRC, "n_src %d outside [1,%d]", n_src, SAI_MAX_SRC);
  int32_t hp[2 + SAI_MAX_SRC] = {ref_pop, tgt_pop};
  for (int k = 0; k < n_src; ++k) {
    SAI_REQUIRE(src_pops[k] >= 0 && src_pops[k] < lay->n_pops, "bad source population index");
    hp[2 + k] = src_pops[k];
  }
  const int n_hp = 2 + n_src;
  const int64_t rows = sai_hist_rows(lay, hp, n_hp);
  const size_t out_n = (size_t)n_src * W * m_max;
  if (int rc = grow(e->hist, sizeof(int32_t) * (size_t)rows * stride + 512)) return rc;
  if (int rc = grow(e->neg, sizeof(int32_t) * 3 * (size_t)n_neg + 256)) return rc;
  if (int rc = grow(e->dd, sizeof(int64_t) * 2 * out_n + 512)) return rc;
  cudaStream_t st = e->s_comp;
  // [hist rows][uint64 missing totals][int32 err] share the hist buffer's tail
  int32_t* d_hist = static_cast<int32_t*>(e->hist.p);
  uint64_t* d_missing = reinterpret_cast<uint64_t*>(
      static_cast<char*>(e->hist.p) + align256(sizeof(int32_t) * (size_t)rows * stride));
  int32_t* d_err = reinterpret_cast<int32_t*>(d_missing + n_hp);
  int32_t* d_ns = static_cast<int32_t*>(e->neg.p);
  int32_t* d_ni = d_ns + n_neg;
  int32_t* d_nv = d_ni + n_neg;
  if (n_neg > 0) {
    SAI_CUDA_CHECK(cudaMemcpyAsync(d_ns, neg_site, sizeof(int32_t) * n_neg, cudaMemcpyHostToDevice, st));
    SAI_CUDA_CHECK(cudaMemcpyAsync(d_ni, neg_ind, sizeof(int32_t) * n_neg, cudaMemcpyHostToDevice, st));
    SAI_CUDA_CHECK(cudaMemcpyAsync(d_nv, neg_val, sizeof(int32_t) * n_neg, cudaMemcpyHostToDevice, st));
  }
  SAI_CUDA_CHECK(cudaMemsetAsync(d_err, 0, sizeof(int32_t), st));
  if (int rc = sai_site_hist(lay, e->packed.p, e->n_sites, hp, n_hp, d_hist, stride, d_missing, st)) return rc;
  int64_t* d_r = static_cast<int64_t*>(e->dd.p);
  int64_t* d_t = d_r + out_n;
  SAI_CUDA_CHECK(cudaMemsetAsync(d_r, 0, sizeof(int64_t) * 2 * out_n, st));
  const int64_t* d_ws = static_cast<const int64_t*>(e->win.p);
  if (int rc = sai_window_dd(lay, e->packed.p, static_cast<const int32_t*>(e->pos.p), e->n_sites, d_ws, d_ws + W,
                             W, d_hist, stride, ref_pop, tgt_pop, src_pops, n_src, neg_off, d_ns, d_ni, d_nv,
                             d_r, d_t, m_max, d_err, st))
    return rc;
  uint64_t h_missing[2 + SAI_MAX_SRC] = {0};
  int32_t h_err = 0;
  SAI_CUDA_CHECK(cudaMemcpyAsync(ref_sum, d_r, sizeof(int64_t) * out_n, cudaMemcpyDeviceToHost, st));
  SAI_CUDA_CHECK(cudaMemcpyAsync(tgt_sum, d_t, sizeof(int64_t) * out_n, cudaMemcpyDeviceToHost, st));
  SAI_CUDA_CHECK(cudaMemcpyAsync(h_missing, d_missing, sizeof(uint64_t) * n_hp, cudaMemcpyDeviceToHost, st));
  SAI_CUDA_CHECK(cudaMemcpyAsync(&h_err, d_err, sizeof(h_err), cudaMemcpyDeviceToHost, st));
  SAI_CUDA_CHECK(cudaStreamSynchronize(st));
  // the bit-planes keep one missing code: every missing call needs its raw value in the table
  for (int q = 0; q < n_hp; ++q) {
    const int64_t have = neg_off[hp[q] + 1] - neg_off[hp[q]];
    SAI_REQUIRE((int64_t)h_missing[q] == have,
                "population %d has %lld missing calls but %lld entries in the negative-value table", hp[q],
                (long long)h_missing[q], (long long)have);
  }
  SAI_REQUIRE(h_err == 0, "a missing source call has no entry in the negative-value table");
  return SAI_OK;
}

int sai_engine_score_resident(sai_engine* e, const sai_job* jobs, int32_t n_jobs, sai_host_results* out) {
  SAI_REQUIRE(e && e->n_jobs > 0, "no previous sai_engine_score_host call");
  const sai_layout* lay = &e->lay;
  if (int rc = validate_jobs(lay, jobs, n_jobs)) return rc;
  if (int rc = check_results(out)) return rc;
  SAI_CUDA_CHECK(cudaSetDevice(e->device));
  const int64_t n_tiles = sai_num_tiles(e->n_sites);
  const int64_t stride = n_tiles * kTile;
  const ResLayout rl = res_layout(e->W, n_jobs);
  if (int rc = grow(e->mask, sizeof(uint32_t) * 2 * (size_t)n_jobs * n_tiles + 256)) return rc;
  if (int rc = grow(e->qval, sizeof(double) * (size_t)n_jobs * stride + 256)) return rc;
  if (int rc = grow(e->res, rl.total + 256)) return rc;
  e->n_jobs = n_jobs;
  for (int j = 0; j < n_jobs; ++j) e->jobs[j] = jobs[j];
  uint32_t* d_mask_u = static_cast<uint32_t*>(e->mask.p);
  if (n_tiles > 0)
    if (int rc = sai_site_flags(lay, e->packed.p, 0, n_tiles, n_tiles, jobs, n_jobs, d_mask_u,
                                d_mask_u + (size_t)n_jobs * n_tiles, static_cast<double*>(e->qval.p), stride, nullptr,
                                nullptr, 0, 0, e->s_comp))
      return rc;
  return run_windows(e, out);
}

int sai_engine_rescore_windows(sai_engine* e, sai_host_results* out) {
  SAI_REQUIRE(e && e->n_jobs > 0, "no previous sai_engine_score_host call");
  if (int rc = check_results(out)) return rc;
  SAI_CUDA_CHECK(cudaSetDevice(e->device));
  return run_windows(e, out);
}

}  // extern "C"
