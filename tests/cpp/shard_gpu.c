/* A plain C host (no torch, no NCCL, no C++) drives a sharded index through the C ABI: 2 id-range shards -- on two
 * GPUs when the box has them, else both on GPU 0 -- fitted by ab_group_fit, searched by ab_group_batch_search, and
 * checked against the exact top-k (ab_bruteforce_topk). This is the call sequence the reference's single-process
 * pybind host would make ("Several GPUs"). */
#include <stdio.h>
#include <stdlib.h>

#include "alaya_b200.h"

static float frand(unsigned *s) {
  *s = *s * 1664525u + 1013904223u;
  return (float)((*s >> 8) & 0xFFFF) / 65536.0f;
}

int main(void) {
  const unsigned n = 6000, d = 32, nq = 40, k = 10, ef = 96, latent = 6;
  unsigned seed = 12345u;
  float *a = malloc(sizeof(float) * latent * d), *base = malloc(sizeof(float) * n * d), *q = malloc(sizeof(float) * nq * d);
  uint32_t *ids = malloc(4 * nq * k), *gt = malloc(4 * nq * k);
  float *dist = malloc(4 * nq * k);
  for (unsigned i = 0; i < latent * d; i++) a[i] = frand(&seed) - 0.5f;
  for (unsigned r = 0; r < n + nq; r++) { /* low-rank rows: z A + noise */
    float z[8];
    float *row = r < n ? base + (size_t)r * d : q + (size_t)(r - n) * d;
    for (unsigned l = 0; l < latent; l++) z[l] = frand(&seed) - 0.5f;
    for (unsigned j = 0; j < d; j++) {
      float v = 0.02f * (frand(&seed) - 0.5f);
      for (unsigned l = 0; l < latent; l++) v += z[l] * a[l * d + j];
      row[j] = v;
    }
  }
  ab_params p = {AB_INDEX_HNSW, AB_QUANT_NONE, AB_METRIC_L2, n, 32, -1};
  int32_t devices[2] = {0, ab_device_count() > 1 ? 1 : 0};
  ab_group *g = NULL;
  if (ab_group_create(&p, devices, 2, 1, &g) != AB_OK) { fprintf(stderr, "create: %s\n", ab_last_error()); return 1; }
  if (ab_group_fit(g, base, n, d, 100) != AB_OK) { fprintf(stderr, "fit: %s\n", ab_last_error()); return 1; }
  for (int rep = 0; rep < 3; rep++) /* consecutive calls alternate the mailbox parities */
    if (ab_group_batch_search(g, q, nq, k, ef, ids, dist) != AB_OK) { fprintf(stderr, "search: %s\n", ab_last_error()); return 1; }
  if (ab_bruteforce_topk(AB_METRIC_L2, base, n, d, q, nq, k, gt, NULL) != AB_OK) { fprintf(stderr, "gt: %s\n", ab_last_error()); return 1; }
  unsigned hits = 0, sorted = 1, second_shard = 0;
  for (unsigned i = 0; i < nq; i++) {
    for (unsigned x = 0; x < k; x++) {
      for (unsigned y = 0; y < k; y++) hits += ids[i * k + x] == gt[i * k + y];
      if (x && dist[i * k + x - 1] > dist[i * k + x]) sorted = 0;
      second_shard += ids[i * k + x] >= n / 2;
    }
  }
  printf("devices=%d,%d recall=%.4f sorted=%u ids_from_second_shard=%u\n", devices[0], devices[1], hits / (double)(nq * k), sorted, second_shard);
  ab_group_destroy(g);
  return 0;
}
