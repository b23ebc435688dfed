/* Host memory bandwidth probe for the reference arm's context (tools/experiments): T pinned threads, each first-touches its
 * own arrays and runs c[i] = a[i] + b[i] (24 B/row, like the Float64 add of the step). gcc -O3 -march=x86-64-v3 -pthread. */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <unistd.h>
static int T;
static size_t N;  /* doubles per thread per array */
static pthread_barrier_t bar;
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void *work(void *arg) {
  long id = (long)arg;
  cpu_set_t set; CPU_ZERO(&set); CPU_SET(id % sysconf(_SC_NPROCESSORS_ONLN), &set);
  pthread_setaffinity_np(pthread_self(), sizeof set, &set);
  double *a = malloc(N * 8), *b = malloc(N * 8), *c = malloc(N * 8);
  for (size_t i = 0; i < N; ++i) { a[i] = i; b[i] = 2.0 * i; c[i] = 0; }
  for (int it = 0; it < 6; ++it) {
    pthread_barrier_wait(&bar);
    for (size_t i = 0; i < N; ++i) c[i] = a[i] + b[i];
    pthread_barrier_wait(&bar);
  }
  volatile double sink = c[N / 2]; (void)sink;
  return NULL;
}
int main(int argc, char **argv) {
  T = argc > 1 ? atoi(argv[1]) : (int)sysconf(_SC_NPROCESSORS_ONLN);
  size_t total = argc > 2 ? (size_t)atoll(argv[2]) : 1000000000ull;
  N = total / T;
  pthread_barrier_init(&bar, NULL, T + 1);
  pthread_t *th = malloc(sizeof(pthread_t) * T);
  for (long i = 0; i < T; ++i) pthread_create(&th[i], NULL, work, (void *)i);
  double best = 1e9;
  for (int it = 0; it < 6; ++it) {
    pthread_barrier_wait(&bar);
    double t0 = now();
    pthread_barrier_wait(&bar);
    double dt = now() - t0;
    if (it && dt < best) best = dt;
  }
  for (int i = 0; i < T; ++i) pthread_join(th[i], NULL);
  printf("threads %d rows %zu: add f64 best %.4f s = %.1f GB/s (24 B/row) = %.0f Mrows/s\n", T, N * T, best, 24.0 * N * T / best / 1e9, N * T / best / 1e6);
  return 0;
}
