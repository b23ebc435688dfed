// refbench.cpp — native harness that times the CPU restatement (oracle.cpp) of the hot-path step
// on the host cores: the `cpu_baseline` / `--impl reference` leg of bench.py.
//
// *** TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT *** (same rule as oracle.cpp).
//
// arrow-rs kernels are single-threaded pure functions (arrow-array/src/array/mod.rs:99); the
// strongest way to run them on a many-core host is what a query engine does: partition the table
// into contiguous row ranges (Arrow "chunks") and run the kernels of each range on its own core.
// This harness does exactly that, natively:
//   * a persistent pool of `threads` workers, each pinned to one CPU of the process's affinity set;
//   * every worker allocates, FIRST-TOUCHES and generates its own row range (so the pages live on
//     its NUMA node), computes the range's take indices and cached null counts (inputs /
//     metadata, outside the timed region), and pre-faults its caller-owned outputs;
//   * a step = barrier, clock, { filter -> take -> add -> sum } per range through the orc_*
//     functions, barrier, clock: the timer is inside C around the barriers, no thread creation,
//     no Python, no allocation in the timed region.
// Step definition = bench.py's GPU step (filter.rs:201, take.rs:89, numeric.rs:36, aggregate.rs:943).
// The per-range results are folded into checksums that bench.py compares with the GPU's.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <time.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/arrow_cuda.h"

extern "C" {
acu_status orc_filter_primitive(const acu_array *predicate, int32_t elem_bytes, const acu_array *values, acu_array_out *out,
                                int64_t *out_count, int32_t *out_strategy);
acu_status orc_take_primitive(int32_t elem_bytes, const acu_array *values, const acu_array *indices, acu_dtype index_dtype,
                              int32_t check_bounds_opt, acu_array_out *out);
acu_status orc_arith(acu_dtype dtype, acu_arith_op op, const acu_array *a, const acu_array *b, acu_array_out *out);
acu_status orc_aggregate(acu_dtype dtype, acu_agg_op op, const acu_array *a, int32_t vector_bytes, uint64_t *out_bits,
                         int64_t *out_valid_count);
acu_status orc_generate_values(int32_t kind, uint64_t seed, int64_t first_row, uint64_t param, void *out, int64_t n);
acu_status orc_generate_bits(uint64_t seed, int64_t first_row, double p, uint8_t *out_bits, int64_t n);
acu_status orc_bitmap_count(const uint8_t *bits, int64_t offset, const uint8_t *validity, int64_t validity_offset, int64_t len,
                            int64_t *out_count);
}

namespace {

double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void *big_alloc(size_t bytes) {
  bytes = (bytes + 4095) & ~(size_t)4095;
  if (bytes == 0) bytes = 4096;
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) return nullptr;
  if (!getenv("ORC_BENCH_NO_THP")) madvise(p, bytes, MADV_HUGEPAGE);
  return p;
}
void big_free(void *p, size_t bytes) {
  bytes = (bytes + 4095) & ~(size_t)4095;
  if (bytes == 0) bytes = 4096;
  if (p) munmap(p, bytes);
}

struct Buf {
  void *p = nullptr;
  size_t bytes = 0;
  bool alloc(size_t b) { bytes = b; p = big_alloc(b); return p != nullptr; }
  void release() { big_free(p, bytes); p = nullptr; }
};

enum Cmd { CMD_NONE = 0, CMD_GENERATE = 1, CMD_STEP = 2, CMD_CHECK = 3, CMD_EXIT = 4 };

struct Part {
  int64_t lo = 0, rows = 0, m = 0;  // global first row, rows, selected rows
  Buf i64, i64_valid, pred, a, b, a_valid, b_valid, idx;
  Buf f_v, f_n, t_v, t_n, s_v, s_n;
  int64_t nc_i64 = 0, nc_a = 0, nc_b = 0;
  // results of the last step
  acu_array_out o_f{}, o_t{}, o_s{};
  uint64_t sum_bits = 0;
  int64_t sum_valid = 0;
  int32_t status = 0;
  // checksums (CMD_CHECK)
  uint64_t chk[8] = {};
};

struct RefBench {
  int64_t rows = 0, first_row = 0;
  int threads = 1;
  uint64_t seeds[7] = {};
  double selectivity = 0.1, null_density = 0.05;
  std::vector<Part> parts;
  std::vector<pthread_t> tids;
  std::vector<int> cpus;
  pthread_barrier_t bar;
  std::atomic<int> cmd{CMD_NONE};
  std::atomic<int> failed{0};
};

struct WorkerArg { RefBench *rb; int k; };

size_t bm_bytes(int64_t n) { return acu_bitmap_bytes(n) + 8; }

acu_array mk(const void *values, const void *validity, int64_t n, int64_t nc) {
  acu_array a{};
  a.values = values;
  a.validity = static_cast<const uint8_t *>(validity);
  a.len = n;
  a.null_count = validity ? nc : 0;
  return a;
}

void generate(RefBench *rb, Part &p) {
  const int64_t n = p.rows, lo = rb->first_row + p.lo;
  bool ok = p.i64.alloc((size_t)n * 8) && p.a.alloc((size_t)n * 8) && p.b.alloc((size_t)n * 8) && p.i64_valid.alloc(bm_bytes(n)) &&
            p.pred.alloc(bm_bytes(n)) && p.a_valid.alloc(bm_bytes(n)) && p.b_valid.alloc(bm_bytes(n));
  if (!ok) { rb->failed = 1; return; }
  // seeds: 0 values(i64), 1 a, 2 b, 3 i64 validity, 4 a validity, 5 b validity, 6 predicate — exactly bench.py's Workload
  orc_generate_values(0, rb->seeds[0], lo, 0, p.i64.p, n);
  orc_generate_values(2, rb->seeds[1], lo, 0, p.a.p, n);
  orc_generate_values(2, rb->seeds[2], lo, 0, p.b.p, n);
  orc_generate_bits(rb->seeds[3], lo, 1.0 - rb->null_density, (uint8_t *)p.i64_valid.p, n);
  orc_generate_bits(rb->seeds[4], lo, 1.0 - rb->null_density, (uint8_t *)p.a_valid.p, n);
  orc_generate_bits(rb->seeds[5], lo, 1.0 - rb->null_density, (uint8_t *)p.b_valid.p, n);
  orc_generate_bits(rb->seeds[6], lo, rb->selectivity, (uint8_t *)p.pred.p, n);
  int64_t c = 0;
  orc_bitmap_count((uint8_t *)p.i64_valid.p, 0, nullptr, 0, n, &c); p.nc_i64 = n - c;
  orc_bitmap_count((uint8_t *)p.a_valid.p, 0, nullptr, 0, n, &c); p.nc_a = n - c;
  orc_bitmap_count((uint8_t *)p.b_valid.p, 0, nullptr, 0, n, &c); p.nc_b = n - c;
  orc_bitmap_count((uint8_t *)p.pred.p, 0, nullptr, 0, n, &c); p.m = c;
  // take indices = the selected rows of the range, ascending, range-local (an INPUT of take)
  if (!p.idx.alloc((size_t)(p.m + 1) * 4)) { rb->failed = 1; return; }
  uint32_t *ix = static_cast<uint32_t *>(p.idx.p);
  const uint8_t *pb = static_cast<const uint8_t *>(p.pred.p);
  int64_t k = 0;
  for (int64_t w = 0; w * 64 < n; ++w) {
    uint64_t bits;
    memcpy(&bits, pb + w * 8, 8);
    if (n - w * 64 < 64) bits &= (~0ull) >> (64 - (n - w * 64));
    while (bits) {
      ix[k++] = (uint32_t)(w * 64 + __builtin_ctzll(bits));
      bits &= bits - 1;
    }
  }
  // caller-owned outputs, pre-faulted (a warm allocator would hand arrow-rs recycled pages)
  ok = p.f_v.alloc((size_t)p.m * 8 + 64) && p.f_n.alloc(bm_bytes(p.m)) && p.t_v.alloc((size_t)p.m * 8 + 64) && p.t_n.alloc(bm_bytes(p.m)) &&
       p.s_v.alloc((size_t)n * 8 + 64) && p.s_n.alloc(bm_bytes(n));
  if (!ok) { rb->failed = 1; return; }
  for (Buf *b : {&p.f_v, &p.f_n, &p.t_v, &p.t_n, &p.s_v, &p.s_n}) memset(b->p, 1, b->bytes);
}

void step(Part &p) {
  const int64_t n = p.rows;
  acu_array pred = mk(p.pred.p, nullptr, n, 0);
  acu_array col = mk(p.i64.p, p.i64_valid.p, n, p.nc_i64);
  acu_array idx = mk(p.idx.p, nullptr, p.m, 0);
  acu_array a = mk(p.a.p, p.a_valid.p, n, p.nc_a);
  acu_array b = mk(p.b.p, p.b_valid.p, n, p.nc_b);
  p.o_f = acu_array_out{}; p.o_f.values = p.f_v.p; p.o_f.validity = (uint8_t *)p.f_n.p;
  p.o_t = acu_array_out{}; p.o_t.values = p.t_v.p; p.o_t.validity = (uint8_t *)p.t_n.p;
  p.o_s = acu_array_out{}; p.o_s.values = p.s_v.p; p.o_s.validity = (uint8_t *)p.s_n.p;
  int64_t cnt = 0;
  int32_t strat = 0;
  acu_status st = orc_filter_primitive(&pred, 8, &col, &p.o_f, &cnt, &strat);
  if (!st) st = orc_take_primitive(8, &col, &idx, ACU_U32, 0, &p.o_t);
  if (!st) st = orc_arith(ACU_F64, ACU_ADD, &a, &b, &p.o_s);
  if (!st) {
    acu_array t = mk(p.o_t.values, p.o_t.has_validity ? p.o_t.validity : nullptr, p.o_t.len, p.o_t.null_count);
    st = orc_aggregate(ACU_I64, ACU_SUM, &t, 16, &p.sum_bits, &p.sum_valid);
  }
  p.status = st;
}

uint64_t wsum(const void *v, int64_t n) {
  const uint64_t *x = static_cast<const uint64_t *>(v);
  uint64_t s = 0;
  for (int64_t i = 0; i < n; ++i) s += x[i];
  return s;
}

// chk: 0 filter rows, 1 filter nulls, 2 wrapping sum of the filtered values (all slots), 3 take nulls,
//      4 wrapping sum of the taken values (all slots), 5 add nulls, 6 wrapping sum of the add output bit patterns, 7 sum(taken) valid rows
void check(Part &p) {
  p.chk[0] = (uint64_t)p.o_f.len;
  p.chk[1] = p.o_f.has_validity ? (uint64_t)p.o_f.null_count : 0;
  p.chk[2] = wsum(p.o_f.values, p.o_f.len);
  p.chk[3] = p.o_t.has_validity ? (uint64_t)p.o_t.null_count : 0;
  p.chk[4] = wsum(p.o_t.values, p.o_t.len);
  p.chk[5] = p.o_s.has_validity ? (uint64_t)p.o_s.null_count : 0;
  p.chk[6] = wsum(p.o_s.values, p.o_s.len);
  p.chk[7] = (uint64_t)p.sum_valid;
}

void *worker(void *argp) {
  WorkerArg *wa = static_cast<WorkerArg *>(argp);
  RefBench *rb = wa->rb;
  const int k = wa->k;
  delete wa;
  if (!rb->cpus.empty()) {
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(rb->cpus[k % rb->cpus.size()], &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
  }
  Part &p = rb->parts[k];
  for (;;) {
    pthread_barrier_wait(&rb->bar);  // A: command published
    const int cmd = rb->cmd.load();
    if (cmd == CMD_GENERATE) generate(rb, p);
    else if (cmd == CMD_STEP) step(p);
    else if (cmd == CMD_CHECK) check(p);
    pthread_barrier_wait(&rb->bar);  // B: command done
    if (cmd == CMD_EXIT) break;
  }
  return nullptr;
}

double run_cmd(RefBench *rb, int cmd) {
  rb->cmd.store(cmd);
  const double t0 = now_s();          // every worker already waits at A: the release is immediate
  pthread_barrier_wait(&rb->bar);     // A
  pthread_barrier_wait(&rb->bar);     // B
  return now_s() - t0;
}

}  // namespace

extern "C" {

// seeds: values(i64), a, b, i64 validity, a validity, b validity, predicate. threads <= 0: one per allowed CPU.
acu_status orc_bench_create(int64_t rows, int64_t first_row, int32_t threads, const uint64_t seeds[7], double selectivity,
                            double null_density, int32_t pin, void **out_handle, double *out_generate_seconds) {
  *out_handle = nullptr;
  RefBench *rb = new RefBench();
  rb->rows = rows;
  rb->first_row = first_row;
  memcpy(rb->seeds, seeds, sizeof(rb->seeds));
  rb->selectivity = selectivity;
  rb->null_density = null_density;
  cpu_set_t set;
  CPU_ZERO(&set);
  std::vector<int> allowed;
  if (sched_getaffinity(0, sizeof(set), &set) == 0)
    for (int c = 0; c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &set)) allowed.push_back(c);
  if (threads <= 0) threads = allowed.empty() ? 1 : (int)allowed.size();
  if ((int64_t)threads > (rows + 63) / 64) threads = (int)((rows + 63) / 64);
  if (threads < 1) threads = 1;
  rb->threads = threads;
  if (pin) {
    rb->cpus = allowed;
    // ORC_BENCH_CPU_ORDER="0,32,1,33,...": the order in which threads are pinned (bench.py passes physical cores first,
    // alternating between the NUMA nodes, so that any thread count spreads over both sockets' memory controllers)
    if (const char *order = getenv("ORC_BENCH_CPU_ORDER")) {
      std::vector<int> ordered;
      for (const char *q = order; *q;) {
        char *end = nullptr;
        const long c = strtol(q, &end, 10);
        if (end == q) break;
        for (int a : allowed)
          if (a == (int)c) { ordered.push_back((int)c); break; }
        q = *end == ',' ? end + 1 : end;
      }
      if (!ordered.empty()) rb->cpus = ordered;
    }
  }
  // contiguous ranges aligned to 64 rows (bitmaps split on u64 words)
  const int64_t per = (((rows + threads - 1) / threads) + 63) / 64 * 64;
  rb->parts.resize(threads);
  for (int k = 0; k < threads; ++k) {
    int64_t lo = (int64_t)k * per;
    if (lo > rows) lo = rows;
    int64_t hi = lo + per;
    if (hi > rows) hi = rows;
    rb->parts[k].lo = lo;
    rb->parts[k].rows = hi - lo;
  }
  pthread_barrier_init(&rb->bar, nullptr, (unsigned)threads + 1);
  rb->tids.resize(threads);
  for (int k = 0; k < threads; ++k) pthread_create(&rb->tids[k], nullptr, worker, new WorkerArg{rb, k});
  const double g = run_cmd(rb, CMD_GENERATE);
  if (out_generate_seconds) *out_generate_seconds = g;
  *out_handle = rb;
  return rb->failed.load() ? ACU_ERR_OUT_OF_MEMORY : ACU_OK;
}

int32_t orc_bench_threads(void *h) { return static_cast<RefBench *>(h)->threads; }

// One timed step over the whole table. out: seconds, the wrapping Int64 sum of the taken rows (bits) and its valid count.
acu_status orc_bench_step(void *h, double *out_seconds, uint64_t *out_sum_bits, int64_t *out_valid) {
  RefBench *rb = static_cast<RefBench *>(h);
  const double s = run_cmd(rb, CMD_STEP);
  uint64_t bits = 0;
  int64_t valid = 0;
  acu_status st = ACU_OK;
  for (Part &p : rb->parts) {
    if (p.status && !st) st = p.status;
    bits += p.sum_bits;  // i64 sum is wrapping: partial sums add up in any order (aggregate.rs:943)
    valid += p.sum_valid;
  }
  if (out_seconds) *out_seconds = s;
  if (out_sum_bits) *out_sum_bits = bits;
  if (out_valid) *out_valid = valid;
  return st;
}

// Checksums of the last step's outputs, folded over the ranges (see check()). Untimed.
acu_status orc_bench_check(void *h, uint64_t out[8]) {
  RefBench *rb = static_cast<RefBench *>(h);
  run_cmd(rb, CMD_CHECK);
  for (int i = 0; i < 8; ++i) out[i] = 0;
  for (Part &p : rb->parts)
    for (int i = 0; i < 8; ++i) out[i] += p.chk[i];
  return ACU_OK;
}

void orc_bench_destroy(void *h) {
  RefBench *rb = static_cast<RefBench *>(h);
  if (!rb) return;
  run_cmd(rb, CMD_EXIT);
  for (pthread_t t : rb->tids) pthread_join(t, nullptr);
  pthread_barrier_destroy(&rb->bar);
  for (Part &p : rb->parts)
    for (Buf *b : {&p.i64, &p.i64_valid, &p.pred, &p.a, &p.b, &p.a_valid, &p.b_valid, &p.idx, &p.f_v, &p.f_n, &p.t_v, &p.t_n, &p.s_v, &p.s_n})
      b->release();
  delete rb;
}

}  // extern "C"
